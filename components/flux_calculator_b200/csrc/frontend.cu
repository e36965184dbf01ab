// frontend.cu -- "next" rows 3 and 4 of the scope table: the reference's own configuration
// files feed the context directly, so that a standalone harness can consume an unchanged flux_calculator.nml and
// the corrections/mass_evap-MM.nc files.  Host code only (no kernels).
//
//   fc_configure_from_namelist : &input which_* method arrays (flux_calculator.F90:99-107, NAMELIST /input/ :109-130)
//                                and &correctionsctl init_date, lcorrections (bias_corrections.F90:60-76)
//   fc_load_corrections        : initialize_bias_corrections (bias_corrections.F90:165-249): 12 monthly files,
//                                variable 'mass_evap', slice (offset, size), _FillValue -> 0, then fc_set_corrections
//   fc_namelist_get, fc_nc_read_var_double : the two parsers on their own (no device needed)
//
// The NetCDF reader handles the classic formats (CDF-1, CDF-2 = 64-bit offset, CDF-5) from their published layout;
// NetCDF-4/HDF5 files are recognised and refused with a message (convert with `nccopy -k classic`, or let the host
// pass the array to fc_set_corrections as before).
#include "context.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

using namespace fc;

namespace {

// ---------------------------------------------------------------------------------------------
// Fortran namelist input (F2003 10.10): &group name[(subscripts)] = value list ... /
// ---------------------------------------------------------------------------------------------
struct NmlValue {
    bool null = true;
    bool quoted = false;
    std::string text;
};
struct NmlSub {
    bool range = false;      // ':' section
    long lo = 0, hi = 0;     // 0 = bound omitted
};
struct NmlAssign {
    std::string name;
    std::vector<NmlSub> subs;
    std::vector<NmlValue> vals;
};
struct NmlGroup {
    std::string name;
    std::vector<NmlAssign> items;
};

std::string lower(std::string s)
{
    for (char &c : s) c = (char)tolower((unsigned char)c);
    return s;
}

struct NmlParser {
    const std::string &t;
    size_t p = 0;
    std::string err;
    explicit NmlParser(const std::string &text) : t(text) {}

    void skip_blank()
    {
        for (;;) {
            while (p < t.size() && isspace((unsigned char)t[p])) ++p;
            if (p < t.size() && t[p] == '!') {
                while (p < t.size() && t[p] != '\n') ++p;
                continue;
            }
            break;
        }
    }
    static bool ident_start(char c) { return isalpha((unsigned char)c) || c == '_'; }
    static bool ident_char(char c) { return isalnum((unsigned char)c) || c == '_'; }

    // at p: identifier [ '(' ... ')' ] '='  ?  (a new assignment begins)
    bool at_assignment() const
    {
        size_t q = p;
        if (q >= t.size() || !ident_start(t[q])) return false;
        while (q < t.size() && ident_char(t[q])) ++q;
        while (q < t.size() && (t[q] == ' ' || t[q] == '\t')) ++q;
        if (q < t.size() && t[q] == '(') {
            int depth = 0;
            while (q < t.size()) {
                if (t[q] == '(') ++depth;
                if (t[q] == ')' && --depth == 0) {
                    ++q;
                    break;
                }
                ++q;
            }
            while (q < t.size() && (t[q] == ' ' || t[q] == '\t')) ++q;
        }
        return q < t.size() && t[q] == '=';
    }

    bool parse_subs(std::vector<NmlSub> &subs)
    {
        ++p;      // '('
        for (;;) {
            NmlSub s;
            auto num = [&](long &v) {
                while (p < t.size() && isspace((unsigned char)t[p])) ++p;
                size_t q = p;
                if (q < t.size() && (t[q] == '-' || t[q] == '+')) ++q;
                while (q < t.size() && isdigit((unsigned char)t[q])) ++q;
                if (q == p) return false;
                v = strtol(t.substr(p, q - p).c_str(), nullptr, 10);
                p = q;
                while (p < t.size() && isspace((unsigned char)t[p])) ++p;
                return true;
            };
            const bool has_lo = num(s.lo);
            while (p < t.size() && isspace((unsigned char)t[p])) ++p;
            if (p < t.size() && t[p] == ':') {
                ++p;
                s.range = true;
                if (!has_lo) s.lo = 0;
                if (!num(s.hi)) s.hi = 0;
            } else if (!has_lo) {
                err = "namelist: bad subscript";
                return false;
            }
            subs.push_back(s);
            while (p < t.size() && isspace((unsigned char)t[p])) ++p;
            if (p < t.size() && t[p] == ',') {
                ++p;
                continue;
            }
            if (p < t.size() && t[p] == ')') {
                ++p;
                return true;
            }
            err = "namelist: unterminated subscript list";
            return false;
        }
    }

    bool parse_values(std::vector<NmlValue> &vals)
    {
        bool pending_sep = true;      // a value may follow (start of list or after a comma)
        for (;;) {
            skip_blank();
            if (p >= t.size()) {
                err = "namelist: group not terminated by '/'";
                return false;
            }
            const char c = t[p];
            if (c == '/' || (c == '&' && lower(t.substr(p, 4)) == "&end") || at_assignment()) return true;
            if (c == ',') {
                if (pending_sep) vals.push_back(NmlValue());      // null value between two separators
                pending_sep = true;
                ++p;
                continue;
            }
            // optional repeat count r*
            long rep = 1;
            {
                size_t q = p;
                while (q < t.size() && isdigit((unsigned char)t[q])) ++q;
                if (q > p && q < t.size() && t[q] == '*') {
                    rep = strtol(t.substr(p, q - p).c_str(), nullptr, 10);
                    p = q + 1;
                }
            }
            NmlValue v;
            if (p < t.size() && (t[p] == '\'' || t[p] == '"')) {
                const char d = t[p++];
                v.null = false;
                v.quoted = true;
                for (;;) {
                    if (p >= t.size()) {
                        err = "namelist: unterminated character constant";
                        return false;
                    }
                    if (t[p] == d) {
                        if (p + 1 < t.size() && t[p + 1] == d) {
                            v.text += d;
                            p += 2;
                            continue;
                        }
                        ++p;
                        break;
                    }
                    v.text += t[p++];
                }
            } else {
                size_t q = p;
                while (q < t.size() && !isspace((unsigned char)t[q]) && t[q] != ',' && t[q] != '/' && t[q] != '!') ++q;
                if (q > p) {
                    v.null = false;
                    v.text = t.substr(p, q - p);
                }
                p = q;      // r* followed by a separator: rep null values
            }
            for (long k = 0; k < rep; ++k) vals.push_back(v);
            pending_sep = false;
        }
    }

    bool parse(std::vector<NmlGroup> &groups)
    {
        for (;;) {
            // text outside groups is ignored (F2003 10.10.1.1: records before the group name are skipped)
            while (p < t.size() && t[p] != '&' && t[p] != '$') {
                if (t[p] == '!')
                    while (p < t.size() && t[p] != '\n') ++p;
                else
                    ++p;
            }
            if (p >= t.size()) return true;
            ++p;
            NmlGroup g;
            while (p < t.size() && ident_char(t[p])) g.name += t[p++];
            g.name = lower(g.name);
            if (g.name == "end") continue;
            for (;;) {
                skip_blank();
                if (p >= t.size()) {
                    err = "namelist: group &" + g.name + " not terminated by '/'";
                    return false;
                }
                if (t[p] == '/') {
                    ++p;
                    break;
                }
                if (t[p] == '&' && lower(t.substr(p, 4)) == "&end") {
                    p += 4;
                    break;
                }
                if (t[p] == ',') {
                    ++p;
                    continue;
                }
                if (!at_assignment()) {
                    err = "namelist: expected 'name =' in group &" + g.name + " near '" + t.substr(p, 20) + "'";
                    return false;
                }
                NmlAssign a;
                while (p < t.size() && ident_char(t[p])) a.name += t[p++];
                a.name = lower(a.name);
                while (p < t.size() && (t[p] == ' ' || t[p] == '\t')) ++p;
                if (t[p] == '(' && !parse_subs(a.subs)) return false;
                while (p < t.size() && (t[p] == ' ' || t[p] == '\t')) ++p;
                ++p;      // '='
                if (!parse_values(a.vals)) return false;
                g.items.push_back(a);
            }
            groups.push_back(g);
        }
    }
};

// values of one namelist array of the given declared shape (column-major), later assignments override earlier ones
bool nml_resolve(const std::vector<NmlGroup> &groups, const std::string &group, const std::string &name,
                 const std::vector<int64_t> &shape, std::vector<NmlValue> &out, std::string &err)
{
    int64_t total = 1;
    for (int64_t e : shape) total *= e;
    out.assign((size_t)total, NmlValue());
    const size_t rank = shape.size();
    for (const NmlGroup &g : groups) {
        if (g.name != group) continue;
        for (const NmlAssign &a : g.items) {
            if (a.name != name) continue;
            std::vector<int64_t> targets;
            if (a.subs.empty()) {
                for (int64_t k = 0; k < total; ++k) targets.push_back(k);
            } else {
                if (a.subs.size() != rank) {
                    err = "namelist: " + name + " has rank " + std::to_string(rank) + " but " + std::to_string(a.subs.size()) + " subscripts";
                    return false;
                }
                bool element = true;
                std::vector<int64_t> lo(rank), hi(rank);
                for (size_t d = 0; d < rank; ++d) {
                    const NmlSub &s = a.subs[d];
                    lo[d] = s.range ? (s.lo ? s.lo : 1) : s.lo;
                    hi[d] = s.range ? (s.hi ? s.hi : shape[d]) : s.lo;
                    element = element && !s.range;
                    if (lo[d] < 1 || hi[d] > shape[d] || lo[d] > hi[d]) {
                        err = "namelist: subscript of " + name + " out of bounds";
                        return false;
                    }
                }
                if (element) {      // a(i,j) = v1, v2, ...: array element order from that element on (common extension)
                    int64_t lin = 0, mul = 1;
                    for (size_t d = 0; d < rank; ++d) {
                        lin += (lo[d] - 1) * mul;
                        mul *= shape[d];
                    }
                    for (int64_t k = lin; k < total; ++k) targets.push_back(k);
                } else {            // section, column-major
                    std::vector<int64_t> idx(lo);
                    for (;;) {
                        int64_t lin = 0, mul = 1;
                        for (size_t d = 0; d < rank; ++d) {
                            lin += (idx[d] - 1) * mul;
                            mul *= shape[d];
                        }
                        targets.push_back(lin);
                        size_t d = 0;
                        while (d < rank && ++idx[d] > hi[d]) {
                            idx[d] = lo[d];
                            ++d;
                        }
                        if (d == rank) break;
                    }
                }
            }
            if (a.vals.size() > targets.size()) {
                err = "namelist: too many values for " + name;
                return false;
            }
            for (size_t k = 0; k < a.vals.size(); ++k)
                if (!a.vals[k].null) out[(size_t)targets[k]] = a.vals[k];
        }
    }
    return true;
}

bool read_file(const char *path, std::string &text)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    text = ss.str();
    return true;
}

bool nml_logical(const std::string &s, bool &v)
{
    std::string u = lower(s);
    if (!u.empty() && u[0] == '.') u = u.substr(1);
    if (u.empty()) return false;
    if (u[0] == 't') v = true;
    else if (u[0] == 'f') v = false;
    else return false;
    return true;
}

// ---------------------------------------------------------------------------------------------
// NetCDF classic reader (CDF-1 / CDF-2 / CDF-5), big endian, header grammar of the NetCDF format specification
// ---------------------------------------------------------------------------------------------
struct NcFile {
    std::vector<unsigned char> buf;      // header only is parsed eagerly; data is read by seek
    FILE *f = nullptr;
    int version = 0;
    size_t pos = 0;
    std::string err;
    ~NcFile()
    {
        if (f) fclose(f);
    }
    bool need(size_t n)
    {
        while (buf.size() < pos + n) {
            unsigned char tmp[65536];
            const size_t got = fread(tmp, 1, sizeof tmp, f);
            if (got == 0) {
                err = "unexpected end of file in the NetCDF header";
                return false;
            }
            buf.insert(buf.end(), tmp, tmp + got);
        }
        return true;
    }
    bool u32(uint64_t &v)
    {
        if (!need(4)) return false;
        v = ((uint64_t)buf[pos] << 24) | ((uint64_t)buf[pos + 1] << 16) | ((uint64_t)buf[pos + 2] << 8) | buf[pos + 3];
        pos += 4;
        return true;
    }
    bool u64(uint64_t &v)
    {
        if (!need(8)) return false;
        v = 0;
        for (int k = 0; k < 8; ++k) v = (v << 8) | buf[pos + k];
        pos += 8;
        return true;
    }
    bool nonneg(uint64_t &v) { return version == 5 ? u64(v) : u32(v); }      // NON_NEG: 4 bytes, 8 in CDF-5
    bool offset(uint64_t &v) { return version == 1 ? u32(v) : u64(v); }
    bool name(std::string &s)
    {
        uint64_t n;
        if (!nonneg(n) || !need((size_t)((n + 3) & ~3ull))) return false;
        s.assign((const char *)&buf[pos], (size_t)n);
        pos += (size_t)((n + 3) & ~3ull);
        return true;
    }
};

size_t nc_type_size(uint64_t t)
{
    switch (t) {
        case 1: case 2: case 7: return 1;      // byte, char, ubyte
        case 3: case 8: return 2;              // short, ushort
        case 4: case 5: case 9: return 4;      // int, float, uint
        case 6: case 10: case 11: return 8;    // double, int64, uint64
        default: return 0;
    }
}

double nc_decode(const unsigned char *q, uint64_t type)
{
    uint64_t raw = 0;
    const size_t n = nc_type_size(type);
    for (size_t k = 0; k < n; ++k) raw = (raw << 8) | q[k];
    switch (type) {
        case 1: return (double)(int8_t)raw;
        case 2: case 7: return (double)(uint8_t)raw;
        case 3: return (double)(int16_t)raw;
        case 8: return (double)(uint16_t)raw;
        case 4: return (double)(int32_t)raw;
        case 9: return (double)(uint32_t)raw;
        case 5: {
            const uint32_t u = (uint32_t)raw;
            float f;
            memcpy(&f, &u, 4);
            return (double)f;
        }
        case 6: {
            double d;
            memcpy(&d, &raw, 8);
            return d;
        }
        case 10: return (double)(int64_t)raw;
        case 11: return (double)raw;
        default: return 0.0;
    }
}

struct NcAtt {
    std::string name;
    uint64_t type = 0, nelems = 0;
    std::vector<unsigned char> raw;
};
struct NcVar {
    std::string name;
    std::vector<uint64_t> dimids;
    std::vector<NcAtt> atts;
    uint64_t type = 0, vsize = 0, begin = 0;
};

bool nc_att_list(NcFile &F, std::vector<NcAtt> &atts)
{
    uint64_t tag, n;
    if (!F.u32(tag) || !F.nonneg(n)) return false;
    if (tag == 0 && n == 0) return true;
    if (tag != 0x0C) {
        F.err = "bad attribute list tag";
        return false;
    }
    for (uint64_t k = 0; k < n; ++k) {
        NcAtt a;
        if (!F.name(a.name) || !F.u32(a.type) || !F.nonneg(a.nelems)) return false;
        const size_t bytes = (size_t)(a.nelems * nc_type_size(a.type));
        const size_t padded = (bytes + 3) & ~(size_t)3;
        if (!F.need(padded)) return false;
        a.raw.assign(F.buf.begin() + F.pos, F.buf.begin() + F.pos + bytes);
        F.pos += padded;
        atts.push_back(a);
    }
    return true;
}

// rc: 0 ok, 1 cannot open, 2 not a classic NetCDF file, 3 variable not found, 4 start/count outside the variable
int nc_read_var(const char *path, const char *varname, int64_t start0, int64_t count, double *out, double *fill, int *has_fill,
                std::string &err)
{
    NcFile F;
    F.f = fopen(path, "rb");
    if (!F.f) {
        err = std::string("cannot open ") + path;
        return 1;
    }
    if (!F.need(4)) {
        err = std::string(path) + ": " + F.err;
        return 2;
    }
    if (memcmp(&F.buf[0], "\x89HDF", 4) == 0) {
        err = std::string(path) + " is a NetCDF-4/HDF5 file; this reader handles the classic formats (nccopy -k classic)";
        return 2;
    }
    if (memcmp(&F.buf[0], "CDF", 3) != 0 || (F.buf[3] != 1 && F.buf[3] != 2 && F.buf[3] != 5)) {
        err = std::string(path) + " is not a NetCDF classic file";
        return 2;
    }
    F.version = F.buf[3];
    F.pos = 4;
    uint64_t numrecs, tag, n;
    if (!F.nonneg(numrecs)) goto bad;
    {
        // dimensions
        std::vector<uint64_t> dimlen;
        if (!F.u32(tag) || !F.nonneg(n)) goto bad;
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0A) goto bad;
            for (uint64_t k = 0; k < n; ++k) {
                std::string nm;
                uint64_t len;
                if (!F.name(nm) || !F.nonneg(len)) goto bad;
                dimlen.push_back(len);
            }
        }
        std::vector<NcAtt> gatts;
        if (!nc_att_list(F, gatts)) goto bad;
        // variables
        std::vector<NcVar> vars;
        if (!F.u32(tag) || !F.nonneg(n)) goto bad;
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0B) goto bad;
            for (uint64_t k = 0; k < n; ++k) {
                NcVar v;
                uint64_t nd;
                if (!F.name(v.name) || !F.nonneg(nd)) goto bad;
                for (uint64_t d = 0; d < nd; ++d) {
                    uint64_t id;
                    if (!F.nonneg(id)) goto bad;
                    v.dimids.push_back(id);
                }
                if (!nc_att_list(F, v.atts) || !F.u32(v.type) || !F.nonneg(v.vsize) || !F.offset(v.begin)) goto bad;
                vars.push_back(v);
            }
        }
        const NcVar *V = nullptr;
        uint64_t recsize = 0;
        for (const NcVar &v : vars) {
            if (v.name == varname) V = &v;
            if (!v.dimids.empty() && v.dimids[0] < dimlen.size() && dimlen[v.dimids[0]] == 0) recsize += v.vsize;
        }
        if (!V) {
            err = std::string("variable ") + varname + " not found in " + path;
            return 3;
        }
        const size_t esz = nc_type_size(V->type);
        if (esz == 0) goto bad;
        const bool is_rec = !V->dimids.empty() && V->dimids[0] < dimlen.size() && dimlen[V->dimids[0]] == 0;
        uint64_t inner = 1, total = 1;
        for (size_t d = 0; d < V->dimids.size(); ++d) {
            if (V->dimids[d] >= dimlen.size()) goto bad;
            const uint64_t len = (d == 0 && is_rec) ? numrecs : dimlen[V->dimids[d]];
            total *= len;
            if (!(d == 0 && is_rec)) inner *= len;
        }
        if (start0 < 0 || count < 0 || (uint64_t)(start0 + count) > total) {
            err = std::string("start/count outside variable ") + varname + " of " + path + " (" + std::to_string(total) + " elements)";
            return 4;
        }
        if (has_fill) *has_fill = 0;
        for (const NcAtt &a : V->atts)
            if (a.name == "_FillValue" && a.nelems >= 1 && a.raw.size() >= nc_type_size(a.type)) {
                if (fill) *fill = nc_decode(a.raw.data(), a.type);
                if (has_fill) *has_fill = 1;
            }
        std::vector<unsigned char> raw;
        if (!is_rec) {
            raw.resize((size_t)count * esz);
            if (fseeko(F.f, (off_t)(V->begin + (uint64_t)start0 * esz), SEEK_SET) != 0 ||
                fread(raw.data(), 1, raw.size(), F.f) != raw.size()) {
                err = std::string("short read of ") + varname + " in " + path;
                return 4;
            }
            for (int64_t k = 0; k < count; ++k) out[k] = nc_decode(&raw[(size_t)k * esz], V->type);
        } else {      // record variable: element e = (record e / inner, e mod inner)
            raw.resize(esz);
            for (int64_t k = 0; k < count; ++k) {
                const uint64_t e = (uint64_t)(start0 + k), r = e / inner, i = e % inner;
                if (fseeko(F.f, (off_t)(V->begin + r * recsize + i * esz), SEEK_SET) != 0 || fread(raw.data(), 1, esz, F.f) != esz) {
                    err = std::string("short read of ") + varname + " in " + path;
                    return 4;
                }
                out[k] = nc_decode(raw.data(), V->type);
            }
        }
        return 0;
    }
bad:
    err = std::string(path) + ": malformed NetCDF classic header" + (F.err.empty() ? "" : (" (" + F.err + ")"));
    return 2;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int fc_namelist_get(const char *path, const char *group, const char *name, const int64_t *shape, int rank,
                               const int64_t *index, char *out, int outlen)
{
    if (!path || !group || !name || rank < 0 || rank > 7 || (rank > 0 && (!shape || !index)) || !out || outlen < 1)
        return fail(nullptr, FC_ERR_ARG, "fc_namelist_get: bad argument");
    std::string text;
    if (!read_file(path, text)) return fail(nullptr, FC_ERR_ARG, "cannot open namelist file %s", path);
    NmlParser P(text);
    std::vector<NmlGroup> groups;
    if (!P.parse(groups)) return fail(nullptr, FC_ERR_ARG, "%s: %s", path, P.err.c_str());
    std::vector<int64_t> shp(shape, shape + rank);
    if (rank == 0) shp.push_back(1);
    std::vector<NmlValue> vals;
    std::string err;
    if (!nml_resolve(groups, lower(group), lower(name), shp, vals, err)) return fail(nullptr, FC_ERR_ARG, "%s: %s", path, err.c_str());
    int64_t lin = 0, mul = 1;
    for (int d = 0; d < rank; ++d) {
        if (index[d] < 1 || index[d] > shape[d]) return fail(nullptr, FC_ERR_ARG, "fc_namelist_get: index out of bounds");
        lin += (index[d] - 1) * mul;
        mul *= shape[d];
    }
    out[0] = 0;
    if (vals[(size_t)lin].null) return FC_NML_UNSET;
    snprintf(out, (size_t)outlen, "%s", vals[(size_t)lin].text.c_str());
    return FC_OK;
}

extern "C" int fc_nc_read_var_double(const char *path, const char *varname, int64_t start0, int64_t count, double *out,
                                     double *fill_value, int *has_fill_value)
{
    if (!path || !varname || !out) return fail(nullptr, FC_ERR_ARG, "fc_nc_read_var_double: NULL argument");
    std::string err;
    const int rc = nc_read_var(path, varname, start0, count, out, fill_value, has_fill_value, err);
    if (rc) return fail(nullptr, FC_ERR_ARG, "%s", err.c_str()), rc + 100;      // 101 open, 102 format, 103 variable, 104 range
    return FC_OK;
}

extern "C" int fc_configure_from_namelist(fc_context *c, const char *path, int bottom_model)
{
    if (!c || !path || bottom_model < 1 || bottom_model > 10) return fail(c, FC_ERR_ARG, "fc_configure_from_namelist: bad argument");
    std::string text;
    if (!read_file(path, text)) return fail(c, FC_ERR_ARG, "cannot open namelist file %s", path);
    NmlParser P(text);
    std::vector<NmlGroup> groups;
    if (!P.parse(groups)) return fail(c, FC_ERR_ARG, "%s: %s", path, P.err.c_str());
    static const char *which[] = {"which_spec_vapor_surface_t", "which_spec_vapor_surface_u", "which_spec_vapor_surface_v",
                                  "which_flux_mass_evap", "which_flux_heat_latent", "which_flux_heat_sensible",
                                  "which_flux_momentum", "which_flux_radiation_blackbody"};
    const std::vector<int64_t> shape = {10, 10};      // (MAX_BOTTOM_MODELS, MAX_SURFACE_TYPES), basic.F90:27-28
    std::string err;
    for (const char *w : which) {
        std::vector<NmlValue> vals;
        if (!nml_resolve(groups, "input", w, shape, vals, err)) return fail(c, FC_ERR_ARG, "%s: %s", path, err.c_str());
        for (int i = 1; i <= c->S; ++i) {
            const NmlValue &v = vals[(size_t)((bottom_model - 1) + 10 * (i - 1))];
            std::string m = v.null ? "none" : v.text;      // declared default 'none' (flux_calculator.F90:99-107)
            while (!m.empty() && m.back() == ' ') m.pop_back();
            if (int rc = fc_set_method(c, w, i, m.c_str())) return rc;
        }
    }
    // &correctionsctl (bias_corrections.F90:60-76); a missing group leaves lcorrections = .FALSE. like the reference
    std::vector<NmlValue> v;
    c->nml_lcorrections = false;
    if (nml_resolve(groups, "correctionsctl", "lcorrections", {1}, v, err) && !v[0].null) {
        bool b = false;
        if (!nml_logical(v[0].text, b)) return fail(c, FC_ERR_ARG, "%s: lcorrections = %s is not a logical", path, v[0].text.c_str());
        c->nml_lcorrections = b;
    }
    if (nml_resolve(groups, "correctionsctl", "init_date", {1}, v, err) && !v[0].null) {
        char *end = nullptr;
        const long d = strtol(v[0].text.c_str(), &end, 10);
        if (end == v[0].text.c_str() || d < 10101 || d > 99991231) return fail(c, FC_ERR_ARG, "%s: init_date = %s is not a yyyymmdd date", path, v[0].text.c_str());
        c->init_date = (int)d;
    }
    c->nml_read = true;
    return FC_OK;
}

extern "C" int fc_load_corrections(fc_context *c, const char *root_dir, int64_t grid_offset, int reference_start_quirk)
{
    if (!c || !root_dir || grid_offset < 0) return fail(c, FC_ERR_ARG, "fc_load_corrections: bad argument");
    c->warning.clear();
    const int64_t n = c->n[1];
    if (!c->nml_lcorrections) return fc_set_corrections(c, 1, nullptr, 0, 0, c->init_date);      // lcorrections = .FALSE.
    std::vector<double> corr((size_t)(12 * n), 0.0), month((size_t)n);
    for (int m = 1; m <= 12; ++m) {
        char file[64];
        snprintf(file, sizeof file, "/corrections/mass_evap-%02d.nc", m);      // bias_corrections.F90:206-208
        const std::string path = std::string(root_dir) + file;
        // the reference hands the 0-based grid_offset to nf90_get_var as its 1-based start (bias_corrections.F90:220-222,
        // SURVEY App. F-8): rank 0 reads nothing (start 0 is invalid), the other ranks read shifted by one cell
        const int64_t start0 = reference_start_quirk ? grid_offset - 1 : grid_offset;
        double fill = 0.0;
        int has_fill = 0;
        std::string err;
        const int rc = nc_read_var(path.c_str(), "mass_evap", start0, n, month.data(), &fill, &has_fill, err);
        if (rc) {      // "... Unset correction." and carry on with zeros (:211-227)
            c->warning += err + ". Unset correction.\n";
            continue;
        }
        if (!has_fill) {      // :229-234
            c->warning += "Could not get fill value of " + path + ". Unset correction.\n";
            continue;
        }
        for (int64_t j = 0; j < n; ++j) corr[(size_t)(j * 12 + (m - 1))] = (month[(size_t)j] == fill) ? 0.0 : month[(size_t)j];   // :242
    }
    return fc_set_corrections(c, 1, corr.data(), n, 1, c->init_date);
}

extern "C" const char *fc_last_warning(const fc_context *c) { return c ? c->warning.c_str() : ""; }
