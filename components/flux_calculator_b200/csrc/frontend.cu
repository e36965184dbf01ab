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
#include <memory>
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

// ---------------------------------------------------------------------------------------------
// "next" row 4, the rest: the whole /input/ namelist drives a context.
//
// fc_create_from_namelist re-does, on top of the C ABI, what the reference's main program does between reading the
// namelist and the time loop (flux_calculator.F90 STEP 1.4 - 1.7, lines 340-768) with the helpers of
// flux_calculator_basic.F90 (allocate_localvar :287-308, init_localvar :312-330, distribute_input_field :334-358,
// add_input_field :128-166, add_output_field :170-283, prepare_regridding :362-459) and flux_calculator_prepare.F90
// (required-input lists INCLUDING their quirks, 'copy' = pointer alias, everything else allocates the result):
// the library then owns the host arrays a Fortran host would ALLOCATE, knows which slots alias which, which are
// %allocated, which fields arrive from the coupler (name, grid, early flag) and which are sent, and what is regridded
// between the grids when.  tests/golden/step_golden.json holds what the reference's own source text builds for eight
// namelists (executed by tests/golden/fortran_interp.py); tests/test_standalone.py compares registry and field lists.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxVarsNml = 100;      // MAX_VARS, basic.F90:30

const char *kNames35[FC_MAX_VARNAMES + 1] = {
    "",     "ALBE", "ALBA", "AMOI", "AMOM", "FARE", "FICE", "PATM", "PSUR", "QATM", "TATM", "TSUR",
    "UATM", "VATM", "U10M", "V10M", "CMOM", "CMOI", "CHEA", "QSUR", "HLAT", "HSEN", "MEVA", "MPRE",
    "MRAI", "MSNO", "RBBR", "RLWD", "RLWU", "RSID", "RSIU", "RSIN", "RSDD", "RSDR", "UMOM", "VMOM"};

std::string rtrim(std::string s)
{
    while (!s.empty() && s.back() == ' ') s.pop_back();
    return s;
}

struct SaBuilder {
    fc::Standalone &R;
    const std::vector<NmlGroup> &groups;
    int m;      // my_bottom_model
    std::string err;
    int err_code = FC_OK;

    // namelist arrays with their declared defaults (flux_calculator.F90:62-107)
    std::vector<NmlValue> get(const char *name, std::vector<int64_t> shape)
    {
        std::vector<NmlValue> v;
        std::string e;
        if (!nml_resolve(groups, "input", name, shape, v, e) && err.empty()) {
            err = e;
            err_code = FC_ERR_ARG;
        }
        return v;
    }
    static std::string str(const NmlValue &v, const char *dflt) { return v.null ? std::string(dflt) : rtrim(v.text); }
    static double num(const NmlValue &v, double dflt)
    {
        if (v.null) return dflt;
        std::string t = v.text;
        for (char &ch : t)
            if (ch == 'd' || ch == 'D') ch = 'e';
        return strtod(t.c_str(), nullptr);
    }
    static bool flag(const NmlValue &v, bool dflt)
    {
        bool b = dflt;
        if (!v.null) nml_logical(v.text, b);
        return b;
    }

    int var_index(const std::string &name) const
    {
        for (int i = 1; i <= FC_MAX_VARNAMES; ++i)
            if (name == kNames35[i]) return i;
        return 0;
    }
    bool stop(int code, const std::string &msg)
    {
        if (err.empty()) {
            err = msg;
            err_code = code;
        }
        return false;
    }
    fc::SaSlot &slot(int i, int g, int idx) { return R.slot[i][g][idx]; }
    bool assoc(int i, int g, int idx) { return slot(i, g, idx).arr >= 0; }

    int new_array(int g, double fill, bool has_fill)
    {
        fc::SaArray a;
        a.grid = g;
        a.fill = fill;
        a.has_fill = has_fill;
        R.arrays.push_back(a);
        return (int)R.arrays.size() - 1;
    }
    // allocate_localvar (basic.F90:287-308)
    bool allocate_localvar(const std::string &name, int i, int g)
    {
        const int idx = var_index(name);
        if (!idx) return stop(FC_ERR_ARG, "Could not allocate local variable " + name + " because flux_calculator does not know this variable.");
        fc::SaSlot &s = slot(i, g, idx);
        s.arr = new_array(g, 0.0, false);
        s.allocated = true;
        s.put[1] = s.put[2] = s.put[3] = false;
        return true;
    }
    // distribute_input_field (basic.F90:334-358)
    void distribute(const std::string &name, int g, int from, int to)
    {
        const int idx = var_index(name);
        for (int j = 1; j <= R.S; ++j)
            if (j == to || (j != from && to == 0)) slot(j, g, idx).arr = slot(from, g, idx).arr;
    }
    // add_input_field (basic.F90:128-166)
    void add_input(const std::string &name, char letter, int i, int g)
    {
        fc::SaField f;
        char buf[16];
        snprintf(buf, sizeof buf, "R%c%s%02d", letter, name.c_str(), i);
        f.name = buf;
        f.grid = g;
        f.type = i;
        f.idx = var_index(name);
        f.early = name == "FARE" || name == "TSUR" || name == "ALBE" || name == "CMOM" || name == "CMOI" || name == "CHEA";   // :153-155
        R.in.push_back(f);
    }

    // one grid's share of STEP 1.4 (flux_calculator.F90:436-476 for t, :477-520 u, :521-563 v)
    bool receive_grid(int g, const char *suffix)
    {
        const std::string sfx(suffix);
        auto nb = get(("name_bottom_var_" + sfx).c_str(), {10, 10, kMaxVarsNml});
        auto vb = get(("val_bottom_var_" + sfx).c_str(), {10, 10, kMaxVarsNml});
        auto na = get(("name_atmos_var_" + sfx).c_str(), {kMaxVarsNml});
        auto va = get(("val_atmos_var_" + sfx).c_str(), {kMaxVarsNml});
        if (!err.empty()) return false;
        for (int i = 1; i <= 10; ++i)
            for (int j = 1; j <= kMaxVarsNml; ++j) {
                const size_t k = (size_t)((m - 1) + 10 * (i - 1) + 100 * (j - 1));
                const std::string name = str(nb[k], "none");
                if (name == "none") continue;
                if (!allocate_localvar(name, i, g)) return false;
                const double val = num(vb[k], -1.0e20);
                if (val > -0.99e20) {      // a constant from the namelist (init_localvar)
                    fc::SaArray &a = R.arrays[(size_t)slot(i, g, var_index(name)).arr];
                    a.fill = val;
                    a.has_fill = a.constant = true;
                } else if (val < -1.99e20) {      // -2.0e20: use the field of surface type 1
                    distribute(name, g, 1, 0);
                } else {
                    add_input(name, R.letter, i, g);
                }
            }
        for (int j = 1; j <= kMaxVarsNml; ++j) {
            const std::string name = str(na[(size_t)j - 1], "none");
            if (name == "none") continue;
            if (!allocate_localvar(name, 0, g)) return false;
            const double val = num(va[(size_t)j - 1], -1.0e20);
            if (val > -0.99e20) {
                fc::SaArray &a = R.arrays[(size_t)slot(0, g, var_index(name)).arr];
                a.fill = val;
                a.has_fill = a.constant = true;
            } else {
                add_input(name, 'A', 0, g);
            }
            distribute(name, g, 0, 0);      // every surface type sees the atmosphere's array
        }
        return true;
    }

    // prepare_regridding (basic.F90:362-459) for variable idx, surface type st (0 = all)
    bool prepare_regridding(int idx, int st, const std::vector<NmlValue> rg[4])
    {
        static const int from[4] = {2, 3, 1, 1}, to[4] = {1, 1, 2, 3};      // u->t, v->t, t->u, t->v: the reference's order
        for (int d = 0; d < 4; ++d)
            for (int j = 1; j <= 10; ++j) {
                if (!(j == st || st == 0)) continue;
                for (int k = 1; k <= kMaxVarsNml; ++k) {
                    const std::string name = str(rg[d][(size_t)((m - 1) + 10 * (j - 1) + 100 * (k - 1))], "none");
                    if (name != kNames35[idx]) continue;
                    fc::SaSlot &dst = slot(j, to[d], idx);
                    if (dst.allocated)
                        return stop(FC_ERR_STATE, std::string("Could not regrid local variable ") + kNames35[idx] +
                                    " as requested in the namelist, because it already exists on that grid.");
                    dst.arr = new_array(to[d], 0.0, false);
                    dst.allocated = true;
                    slot(j, from[d], idx).put[to[d]] = true;
                    fc::SaRegrid r{j, idx, from[d], to[d]};
                    R.regrid.push_back(r);
                }
            }
        return true;
    }

    // the prepare_* routines (flux_calculator_prepare.F90): {tested variable, label in the message} per method, quirks kept
    struct Need { int tested; const char *label; };
    bool prepare(const char *var, int out_idx, int i, int g, const std::string &method, const std::vector<Need> &needs, bool known,
                 int copy_tested)
    {
        if (method == "none") return true;
        std::string missing;
        if (method == "copy") {
            if (!assoc(1, g, copy_tested)) missing = std::string(var) + " for surface_type=1 ";
        } else if (method == "zero" && out_idx != FC_QSUR) {
        } else if (known) {
            for (const Need &n : needs)
                if (!assoc(i, g, n.tested)) missing += std::string(" ") + n.label;
        } else {
            char buf[256];
            snprintf(buf, sizeof buf, "Error calculating %s for surface_type %d on the grid %s: Method %s is not known.", var, i,
                     g == 1 ? "t_grid" : (g == 2 ? "u_grid" : "v_grid"), method.c_str());
            return stop(FC_ERR_METHOD, buf);
        }
        if (!rtrim(missing).empty() || (!missing.empty() && method == "copy")) {
            char buf[512];
            snprintf(buf, sizeof buf, "Error calculating %s for surface_type %d on the grid %s: For method %s we are lacking the following variables: %s",
                     var, i, g == 1 ? "t_grid" : (g == 2 ? "u_grid" : "v_grid"), method.c_str(), rtrim(missing).c_str());
            return stop(FC_ERR_MISSING, buf);
        }
        fc::SaSlot &o = slot(i, g, out_idx);      // do_prepare_calculation (prepare.F90:19-44)
        if (method == "copy") {
            o.arr = slot(1, g, out_idx).arr;
        } else {
            o.arr = new_array(g, 0.0, false);
            o.allocated = true;
        }
        return true;
    }

    bool prepare_all(const std::vector<NmlValue> rg[4])
    {
        static const char *wnames[8] = {"which_spec_vapor_surface_t", "which_spec_vapor_surface_u", "which_spec_vapor_surface_v",
                                        "which_flux_mass_evap", "which_flux_heat_latent", "which_flux_heat_sensible",
                                        "which_flux_momentum", "which_flux_radiation_blackbody"};
        std::vector<NmlValue> w[8];
        for (int q = 0; q < 8; ++q) w[q] = get(wnames[q], {10, 10});
        if (!err.empty()) return false;
        auto meth = [&](int q, int i) {
            const std::string s = str(w[q][(size_t)((m - 1) + 10 * (i - 1))], "none");
            R.method[q][i] = s;
            return s;
        };
        const int A = FC_AMOI, P = FC_PSUR, QA = FC_QATM, QS = FC_QSUR, TA = FC_TATM, TS = FC_TSUR, U = FC_UATM, V = FC_VATM;
        for (int i = 1; i <= R.S; ++i)      // flux_calculator.F90:594-598
            for (int g = 1; g <= 3; ++g) {
                const std::string me = meth(g - 1, i);
                if (!prepare("QSUR", FC_QSUR, i, g, me, {{FC_FICE, "FICE"}, {P, "PSUR"}, {TS, "TSUR"}}, me == "CCLM", FC_QSUR)) return false;
            }
        if (!prepare_regridding(FC_QSUR, 0, rg)) return false;
        for (int i = 1; i <= R.S; ++i) {      // :603-605
            const std::string me = meth(3, i);
            std::vector<Need> n;
            if (me == "CCLM") n = {{A, "AMOI"}, {P, "PSUR"}, {QA, "QATM"}, {QS, "QSUR"}, {TA, "TATM"}, {U, "UATM"}, {V, "VATM"}};
            if (me == "MOM5") n = {{FC_CMOI, "CMOI"}, {P, "PSUR"}, {QA, "QATM"}, {QS, "QSUR"}, {TA, "TATM"}, {U, "UATM"}, {V, "VATM"}};
            if (me == "RCO") n = {{QA, "QATM"}, {QS, "TSUR"}, {U, "UATM"}, {V, "VATM"}};      // prepare.F90:107 tests QSUR, says TSUR
            if (!prepare("MEVA", FC_MEVA, i, 1, me, n, me == "CCLM" || me == "MOM5" || me == "RCO", FC_MEVA)) return false;
        }
        if (!prepare_regridding(FC_MEVA, 0, rg)) return false;
        for (int i = 1; i <= R.S; ++i) {      // :610-612; 'copy' tests HSEN of type 1 (prepare.F90:132)
            const std::string me = meth(4, i);
            if (!prepare("HLAT", FC_HLAT, i, 1, me, {{FC_MEVA, "MEVA"}}, me == "water" || me == "ice", FC_HSEN)) return false;
        }
        if (!prepare_regridding(FC_HLAT, 0, rg)) return false;
        for (int i = 1; i <= R.S; ++i) {      // :615-617; TATM is tested where the message says TSUR (prepare.F90:168,178,184)
            const std::string me = meth(5, i);
            std::vector<Need> n;
            if (me == "CCLM") n = {{A, "AMOI"}, {FC_PATM, "PATM"}, {P, "PSUR"}, {QS, "QSUR"}, {TA, "TATM"}, {TA, "TSUR"}, {U, "UATM"}, {V, "VATM"}};
            if (me == "MOM5") n = {{FC_CHEA, "CHEA"}, {FC_PATM, "PATM"}, {P, "PSUR"}, {QS, "QSUR"}, {TA, "TATM"}, {TA, "TSUR"}, {U, "UATM"}, {V, "VATM"}};
            if (me == "RCO") n = {{TA, "TATM"}, {TA, "TSUR"}, {U, "UATM"}, {V, "VATM"}};
            if (!prepare("HSEN", FC_HSEN, i, 1, me, n, me == "CCLM" || me == "MOM5" || me == "RCO", FC_HSEN)) return false;
        }
        if (!prepare_regridding(FC_HSEN, 0, rg)) return false;
        for (int i = 1; i <= R.S; ++i) {      // :622-624
            const std::string me = meth(7, i);
            if (!prepare("RBBR", FC_RBBR, i, 1, me, {{TS, "TSUR"}}, me == "StBo", FC_RBBR)) return false;
        }
        if (!prepare_regridding(FC_RBBR, 0, rg)) return false;
        for (int north = 0; north < 2; ++north) {      // :634-643; UATM is tested where the message of VMOM says VATM (prepare.F90:258,266)
            const int g = north ? 3 : 2, out = north ? FC_VMOM : FC_UMOM;
            const char *wl = north ? "VATM" : "UATM";
            for (int i = 1; i <= R.S; ++i) {
                const std::string me = meth(6, i);
                std::vector<Need> n;
                if (me == "CCLM") n = {{FC_AMOM, "AMOM"}, {P, "PSUR"}, {QS, "QSUR"}, {TA, "TATM"}, {TA, "TSUR"}, {U, wl}};
                if (me == "MOM5") n = {{FC_CMOM, "CMOM"}, {P, "PSUR"}, {QS, "QSUR"}, {TA, "TATM"}, {TA, "TSUR"}, {U, wl}};
                if (me == "RCO") n = {{U, "UATM"}, {V, "VATM"}};
                if (!prepare(north ? "VMOM" : "UMOM", out, i, g, me, n, me == "CCLM" || me == "MOM5" || me == "RCO", out)) return false;
            }
            if (!prepare_regridding(out, 0, rg)) return false;
        }
        return true;
    }

    // add_output_field (basic.F90:170-283)
    bool add_output(const std::string &name, char letter, int st, int g, bool uniform, double dflt)
    {
        char buf[16];
        snprintf(buf, sizeof buf, "R%c%s%02d", letter, name.c_str(), st);
        for (const fc::SaField &f : R.in)
            if (f.name == buf) return true;      // comes in as an input: nothing to send (:195-199)
        const int i = var_index(name);
        if (!i) return stop(FC_ERR_ARG, "Could not add output field for variable " + name + " because flux_calculator does not know this variable.");
        auto give_default = [&](int t) {
            fc::SaSlot &s = slot(t, g, i);
            s.arr = new_array(g, dflt, true);
            s.allocated = true;
            char w[160];
            snprintf(w, sizeof w, "WARNING: Flux %s cannot be calculated for surface_type=%d => set to %g\n", name.c_str(), t, dflt);
            R.warnings += w;
        };
        if (st == 0) {
            if (uniform) {
                for (int j = 1; j <= R.S; ++j)
                    if (assoc(j, g, i) && !assoc(0, g, i)) slot(0, g, i).arr = slot(j, g, i).arr;      // pointer, not %allocated
            } else {
                bool fluxes = true, areas = true;
                for (int j = 1; j <= R.S; ++j) {
                    fluxes = fluxes && assoc(j, g, i);
                    areas = areas && assoc(j, g, FC_FARE);
                }
                if (!areas)
                    return stop(FC_ERR_MISSING, "ERROR: Output field " + name + " has not been defined as uniform (flux_?_uniform=.FALSE.). To calculate its "
                                "average value across different surface_types, their fractional area (FARE) must be given but is missing.");
                if (fluxes && !assoc(0, g, i)) {
                    slot(0, g, i).arr = new_array(g, 0.0, false);
                    slot(0, g, i).allocated = true;
                }
            }
            if (!assoc(0, g, i)) give_default(0);
        } else {
            if (uniform && !assoc(st, g, i))
                for (int j = 1; j <= R.S; ++j)
                    if (assoc(j, g, i) && !assoc(0, g, i)) slot(0, g, i).arr = slot(j, g, i).arr;
            if (!assoc(st, g, i)) give_default(st);
        }
        fc::SaField f;
        snprintf(buf, sizeof buf, "S%c%s%02d", letter, name.c_str(), st);
        f.name = buf;
        f.grid = g;
        f.type = st;
        f.idx = i;
        f.early = name == "RBBR" || name == "TSUR" || name == "FICE" || name == "ALBE";      // :271-273
        R.out.push_back(f);
        return true;
    }

    bool send_all()
    {
        std::vector<NmlValue> ns[3], sa[3], sb[3], su[3], vf[3];
        static const char *sfx[3] = {"t", "u", "v"};
        for (int g = 0; g < 3; ++g) {
            const std::string s(sfx[g]);
            ns[g] = get(("name_send_" + s).c_str(), {kMaxVarsNml});
            sa[g] = get(("send_to_atmos_" + s).c_str(), {kMaxVarsNml});
            sb[g] = get(("send_to_bottom_" + s).c_str(), {10, kMaxVarsNml});
            su[g] = get(("send_uniform_" + s).c_str(), {10, kMaxVarsNml});
            vf[g] = get(("val_flux_" + s).c_str(), {kMaxVarsNml});
        }
        if (!err.empty()) return false;
        // the reference sizes output_field from this count (flux_calculator.F90:651-683) and then adds without checking:
        // more additions than counted is an out-of-bounds write there, refused here
        size_t counted = 0;
        for (int g = 0; g < 3; ++g)
            for (int j = 1; j <= kMaxVarsNml; ++j)
                if (str(ns[g][(size_t)j - 1], "none") != "none")
                    counted += flag(su[g][(size_t)((m - 1) + 10 * (j - 1))], false) ? 1 : (size_t)(1 + R.S);
        for (int j = 1; j <= kMaxVarsNml; ++j)      // :690-761: t, u, v interleaved per j
            for (int g = 0; g < 3; ++g) {
                const std::string name = str(ns[g][(size_t)j - 1], "none");
                if (name == "none") continue;
                const bool to_atmos = flag(sa[g][(size_t)j - 1], true), to_bottom = flag(sb[g][(size_t)((m - 1) + 10 * (j - 1))], true);
                const bool uni = flag(su[g][(size_t)((m - 1) + 10 * (j - 1))], false);
                const double dflt = num(vf[g][(size_t)j - 1], 0.0);
                if (to_atmos) {      // (on the t grid a flux that goes to the atmosphere only is never uniform, on u / v always: :697-700, :718-721)
                    if (!add_output(name, 'A', 0, g + 1, to_bottom ? uni : (g != 0), dflt)) return false;
                }
                if (to_bottom) {
                    if (uni) {
                        if (!add_output(name, R.letter, 1, g + 1, true, dflt)) return false;
                    } else {
                        for (int i = 1; i <= R.S; ++i)
                            if (!add_output(name, R.letter, i, g + 1, false, dflt)) return false;
                    }
                }
            }
        if (R.out.size() > counted)
            return stop(FC_ERR_STATE, "the namelist sends more fields than the reference allocates room for (a uniform flux that goes to the atmosphere "
                        "AND to the bottom model is counted once, flux_calculator.F90:655-660, but added twice, :694-712): undefined in the reference");
        return true;
    }

    bool build(const int64_t grid_size[3])
    {
        for (int g = 0; g < 3; ++g) R.n[g + 1] = grid_size[g];
        auto letters = get("letter_bottom_model", {10});
        R.letter = 'M';
        if (!letters.empty() && !letters[(size_t)m - 1].null && !letters[(size_t)m - 1].text.empty()) R.letter = letters[(size_t)m - 1].text[0];
        // num_surface_types (:362-410): the highest surface type for which a bottom variable is named on any grid
        R.S = 0;
        for (const char *s : {"t", "u", "v"}) {
            auto nb = get((std::string("name_bottom_var_") + s).c_str(), {10, 10, kMaxVarsNml});
            if (!err.empty()) return false;
            for (int i = 1; i <= 10; ++i)
                for (int j = 1; j <= kMaxVarsNml; ++j)
                    if (str(nb[(size_t)((m - 1) + 10 * (i - 1) + 100 * (j - 1))], "none") != "none") R.S = std::max(R.S, i);
        }
        if (R.S < 1) return stop(FC_ERR_ARG, "the namelist names no bottom-model variable for this bottom model: num_surface_types = 0");
        if (!receive_grid(1, "t") || !receive_grid(2, "u") || !receive_grid(3, "v")) return false;
        std::vector<NmlValue> rg[4] = {get("regrid_u_to_t", {10, 10, kMaxVarsNml}), get("regrid_v_to_t", {10, 10, kMaxVarsNml}),
                                       get("regrid_t_to_u", {10, 10, kMaxVarsNml}), get("regrid_t_to_v", {10, 10, kMaxVarsNml})};
        if (!err.empty()) return false;
        const std::vector<fc::SaField> inputs = R.in;
        for (const fc::SaField &f : inputs)      // STEP 1.5 (:575-578)
            if (!prepare_regridding(f.idx, f.type, rg)) return false;
        R.n_input_regrid = R.regrid.size();
        if (!prepare_all(rg)) return false;      // STEP 1.6
        return send_all();                       // STEP 1.7
    }
};

bool parse_namelist_file(const char *path, std::vector<NmlGroup> &groups, std::string &err)
{
    std::string text;
    if (!read_file(path, text)) {
        err = std::string("cannot open namelist file ") + path;
        return false;
    }
    NmlParser P(text);
    if (!P.parse(groups)) {
        err = std::string(path) + ": " + P.err;
        return false;
    }
    return true;
}

std::string registry_json(const fc::Standalone &R)
{
    std::ostringstream o;
    o << "{\"num_surface_types\":" << R.S << ",\"registry\":[";
    bool first = true;
    for (int i = 0; i <= 10; ++i)
        for (int g = 1; g <= 3; ++g)
            for (int k = 1; k <= FC_MAX_VARNAMES; ++k) {
                const fc::SaSlot &s = R.slot[i][g][k];
                if (s.arr < 0) continue;
                const fc::SaArray &a = R.arrays[(size_t)s.arr];
                o << (first ? "" : ",") << "{\"type\":" << i << ",\"grid\":" << g << ",\"var\":\"" << kNames35[k] << "\",\"storage\":" << s.arr
                  << ",\"allocated\":" << (s.allocated ? "true" : "false") << ",\"fill\":";
                if (a.has_fill) {
                    char b[64];
                    snprintf(b, sizeof b, "%.17g", a.fill);
                    o << b;
                } else {
                    o << "null";
                }
                o << ",\"regrid_to\":[";
                bool f2 = true;
                for (int t = 1; t <= 3; ++t)
                    if (s.put[t]) {
                        o << (f2 ? "" : ",") << t;
                        f2 = false;
                    }
                o << "]}";
                first = false;
            }
    auto list = [&](const char *key, const std::vector<fc::SaField> &v) {
        o << "],\"" << key << "\":[";
        for (size_t k = 0; k < v.size(); ++k)
            o << (k ? "," : "") << "{\"name\":\"" << v[k].name << "\",\"grid\":" << v[k].grid << ",\"early\":" << (v[k].early ? "true" : "false")
              << ",\"type\":" << v[k].type << ",\"var\":\"" << kNames35[v[k].idx] << "\"}";
    };
    list("input_fields", R.in);
    list("output_fields", R.out);
    o << "]}";
    return o.str();
}

}  // namespace

// the registry the namelist describes, as JSON text (no device needed): what fc_create_from_namelist will build
extern "C" int fc_namelist_registry(const char *nml_path, int bottom_model, const int64_t grid_size[3], char *out, int64_t outlen)
{
    if (!nml_path || !grid_size || !out || outlen < 2 || bottom_model < 1 || bottom_model > 10)
        return fail(nullptr, FC_ERR_ARG, "fc_namelist_registry: bad argument");
    std::vector<NmlGroup> groups;
    std::string err;
    if (!parse_namelist_file(nml_path, groups, err)) return fail(nullptr, FC_ERR_ARG, "%s", err.c_str());
    fc::Standalone R;
    SaBuilder B{R, groups, bottom_model};
    if (!B.build(grid_size)) return fail(nullptr, B.err_code ? B.err_code : FC_ERR_ARG, "%s", B.err.c_str());
    const std::string js = registry_json(R);
    if ((int64_t)js.size() + 1 > outlen) return fail(nullptr, FC_ERR_NOMEM, "fc_namelist_registry: need %lld bytes", (long long)js.size() + 1);
    memcpy(out, js.c_str(), js.size() + 1);
    return FC_OK;
}

extern "C" int fc_create_from_namelist(fc_context **out, const char *nml_path, int bottom_model, const int64_t grid_size[3], int device)
{
    if (!out || !nml_path || !grid_size || bottom_model < 1 || bottom_model > 10)
        return fail(nullptr, FC_ERR_ARG, "fc_create_from_namelist: bad argument");
    std::vector<NmlGroup> groups;
    std::string err;
    if (!parse_namelist_file(nml_path, groups, err)) return fail(nullptr, FC_ERR_ARG, "%s", err.c_str());
    std::unique_ptr<fc::Standalone> R(new fc::Standalone());
    SaBuilder B{*R, groups, bottom_model};
    if (!B.build(grid_size)) return fail(nullptr, B.err_code ? B.err_code : FC_ERR_ARG, "%s", B.err.c_str());
    fc_context *c = nullptr;
    if (int rc = fc_create(&c, grid_size, R->S, device)) return rc;
    auto bail = [&](int rc) {
        const std::string msg = c->err;
        fc::standalone_free(*R);
        fc_destroy(c);
        return fail(nullptr, rc, "%s", msg.c_str());
    };
    // the arrays a Fortran host would ALLOCATE: page-locked, so that the per-step copies run at link speed
    for (fc::SaArray &a : R->arrays) {
        const int64_t n = std::max<int64_t>(R->n[a.grid], 1);
        void *p = nullptr;
        if (cudaHostAlloc(&p, (size_t)n * sizeof(double), cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            c->err = "fc_create_from_namelist: cudaHostAlloc failed";
            return bail(FC_ERR_NOMEM);
        }
        a.host = (double *)p;
        const double v = a.has_fill ? a.fill : nan("");      // what nobody has written yet reads as NaN, not as garbage
        for (int64_t k = 0; k < n; ++k) a.host[k] = v;
    }
    for (int i = 0; i <= 10; ++i)
        for (int g = 1; g <= 3; ++g)
            for (int k = 1; k <= FC_MAX_VARNAMES; ++k) {
                const fc::SaSlot &s = R->slot[i][g][k];
                if (s.arr < 0 || i > R->S) continue;
                if (int rc = fc_bind_field(c, i, g, k, R->arrays[(size_t)s.arr].host, R->n[g])) return bail(rc);
                if (int rc = fc_set_allocated(c, i, g, k, s.allocated ? 1 : 0)) return bail(rc);
                const fc::SaArray &a = R->arrays[(size_t)s.arr];
                if (a.constant || (a.has_fill && !a.constant && false))
                    if (int rc = fc_mark_static(c, i, g, k, 1)) return bail(rc);
            }
    static const char *wnames[8] = {"which_spec_vapor_surface_t", "which_spec_vapor_surface_u", "which_spec_vapor_surface_v",
                                    "which_flux_mass_evap", "which_flux_heat_latent", "which_flux_heat_sensible",
                                    "which_flux_momentum", "which_flux_radiation_blackbody"};
    for (int q = 0; q < 8; ++q)
        for (int i = 1; i <= R->S; ++i)
            if (int rc = fc_set_method(c, wnames[q], i, R->method[q][i].c_str())) return bail(rc);
    for (const fc::SaField &f : R->out)
        if (int rc = fc_add_output_field(c, f.type, f.grid, f.idx)) return bail(rc);
    c->warning = R->warnings;
    c->sa = R.release();
    if (int rc = fc_configure_from_namelist(c, nml_path, bottom_model)) {      // &correctionsctl (the methods are set already)
        fc_context *dead = c;
        const std::string msg = c->err;
        fc_destroy(dead);
        return fail(nullptr, rc, "%s", msg.c_str());
    }
    *out = c;
    return FC_OK;
}

static int sa_field(fc_context *c, bool output, int j, char name[16], int *grid, int *early, int *surface_type, int *var_idx, double **field,
                    int64_t *n)
{
    if (!c || !c->sa) return fail(c, FC_ERR_STATE, "this context was not created by fc_create_from_namelist");
    const std::vector<fc::SaField> &v = output ? c->sa->out : c->sa->in;
    if (j < 0 || j >= (int)v.size()) return fail(c, FC_ERR_ARG, "field number %d out of range 0..%d", j, (int)v.size() - 1);
    const fc::SaField &f = v[(size_t)j];
    if (name) snprintf(name, 16, "%s", f.name.c_str());
    if (grid) *grid = f.grid;
    if (early) *early = f.early ? 1 : 0;
    if (surface_type) *surface_type = f.type;
    if (var_idx) *var_idx = f.idx;
    const fc::SaSlot &s = c->sa->slot[f.type][f.grid][f.idx];
    if (field) *field = s.arr >= 0 ? c->sa->arrays[(size_t)s.arr].host : nullptr;
    if (n) *n = c->sa->n[f.grid];
    return FC_OK;
}

extern "C" int fc_num_input_fields(const fc_context *c) { return (c && c->sa) ? (int)c->sa->in.size() : -1; }
extern "C" int fc_num_output_fields(const fc_context *c) { return (c && c->sa) ? (int)c->sa->out.size() : -1; }
extern "C" int fc_input_field(fc_context *c, int j, char name[16], int *grid, int *early, int *surface_type, int *var_idx, double **field, int64_t *n)
{
    return sa_field(c, false, j, name, grid, early, surface_type, var_idx, field, n);
}
extern "C" int fc_output_field(fc_context *c, int j, char name[16], int *grid, int *early, int *surface_type, int *var_idx, double **field, int64_t *n)
{
    return sa_field(c, true, j, name, grid, early, surface_type, var_idx, field, n);
}
extern "C" int fc_field_pointer(fc_context *c, int surface_type, int grid, int var_idx, double **field, int64_t *n)
{
    if (!c || !c->sa) return fail(c, FC_ERR_STATE, "this context was not created by fc_create_from_namelist");
    if (surface_type < 0 || surface_type > 10 || grid < 1 || grid > 3 || var_idx < 1 || var_idx > FC_MAX_VARNAMES || !field)
        return fail(c, FC_ERR_ARG, "fc_field_pointer: bad argument");
    const fc::SaSlot &s = c->sa->slot[surface_type][grid][var_idx];
    *field = s.arr >= 0 ? c->sa->arrays[(size_t)s.arr].host : nullptr;
    if (n) *n = c->sa->n[grid];
    return FC_OK;
}

namespace fc {
void standalone_free(Standalone &R)
{
    for (SaArray &a : R.arrays)
        if (a.host) {
            cudaFreeHost(a.host);
            a.host = nullptr;
        }
}
}  // namespace fc
