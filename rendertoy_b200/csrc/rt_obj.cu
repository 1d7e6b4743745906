// rt_obj.cu -- Wavefront OBJ parser behind rendering.load_obj (host code only, no device work).
//
// The reference delegates parsing to pywavefront (rendering/_loaders.py:1, 8-15, 22: pure Python, seconds to minutes
// for a dragon-class file) and consumes, per mesh, the FIRST material's interleaved, face-corner-expanded vertex list.
// This is that consumer's view of the file, built in one pass over the mapped bytes:
//   v / vn / vt            attribute tables (extra components ignored; vt's v defaults to 0)
//   o                      starts a mesh; usemtl selects (or creates) a material by name, shared between meshes
//   f                      corners v, v/t, v//n, v/t/n, 1-based or negative (relative to the table so far); polygons are
//                          fan-triangulated as (c0, c[i-1], c[i]); the corners go to the CURRENT material's list, its
//                          vertex format (has texture / has normal) is fixed by the first corner it ever receives
//   anything else          ignored (g, s, mtllib, l, comments)
// rendering/_loaders.py keeps the line-by-line Python formulation of exactly these rules as the definition this parser is
// tested against (tests/test_host_api.py), and does the scatter's position normalisation (_loaders.py:34-38) in NumPy.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "rt_common.cuh"

namespace {

struct ObjMaterial {
    int format = -1;              // -1: no corner yet; bit 0: T2F, bit 1: N3F (V3F always)
    std::vector<int64_t> corners; // (vi, ti, ni) per emitted corner, 0-based, -1 when absent
};

struct ObjMesh {
    std::vector<int> materials; // indices into ObjFile::materials, in order of first use
    int64_t n_faces = 0;        // triangles after fan triangulation
};

struct ObjFile {
    std::vector<float> pos, nrm, tex; // xyz, xyz, uv
    std::vector<ObjMaterial> materials;
    std::unordered_map<std::string, int> by_name;
    std::vector<ObjMesh> meshes;
    std::vector<int> listed; // meshes that have a material (what load_obj returns), in file order
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f'; }

const double POW10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                          1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// Decimal text -> double, correctly rounded.  Fast path (Clinger): at most 15 significant digits and |exponent| <= 22 --
// mantissa and power of ten are both exact doubles, so one multiplication or division rounds once.  Everything else
// (long mantissas, big exponents, inf / nan spellings) goes to strtod on a NUL-terminated copy.
bool parse_double(const char *p, const char *end, double *out)
{
    const char *s = p;
    bool neg = false;
    if (s < end && (*s == '+' || *s == '-')) { neg = *s == '-'; ++s; }
    uint64_t mant = 0;
    int digits = 0, exp10 = 0;
    bool any = false, fast = true;
    while (s < end && *s >= '0' && *s <= '9') {
        any = true;
        if (mant != 0 || *s != '0') { if (digits < 19) { mant = mant * 10 + (uint64_t)(*s - '0'); ++digits; } else { fast = false; } }
        ++s;
    }
    if (s < end && *s == '.') {
        ++s;
        while (s < end && *s >= '0' && *s <= '9') {
            any = true;
            if (mant != 0 || *s != '0') { if (digits < 19) { mant = mant * 10 + (uint64_t)(*s - '0'); ++digits; --exp10; } else { fast = false; } }
            else --exp10;
            ++s;
        }
    }
    if (any && s < end && (*s == 'e' || *s == 'E')) {
        const char *t = s + 1;
        bool eneg = false;
        if (t < end && (*t == '+' || *t == '-')) { eneg = *t == '-'; ++t; }
        if (t < end && *t >= '0' && *t <= '9') {
            int e = 0;
            while (t < end && *t >= '0' && *t <= '9') { if (e < 100000) e = e * 10 + (*t - '0'); ++t; }
            exp10 += eneg ? -e : e;
            s = t;
        } else {
            any = false; // "1e" is not a number
        }
    }
    if (any && s == end && fast && digits <= 15 && exp10 >= -22 && exp10 <= 22) {
        double v = (double)mant;
        v = exp10 < 0 ? v / POW10[-exp10] : v * POW10[exp10];
        *out = neg ? -v : v;
        return true;
    }
    // slow path, and the verdict on malformed tokens
    std::string tmp(p, end);
    char *stop = nullptr;
    errno = 0;
    const double v = strtod(tmp.c_str(), &stop);
    if (stop == tmp.c_str() || *stop != '\0') return false;
    if (tmp.find_first_of("xXpP") != std::string::npos) return false; // hex floats: not numbers in an OBJ (nor for Python's float())
    *out = v;
    return true;
}

bool parse_int(const char *p, const char *end, int64_t *out)
{
    const char *s = p;
    bool neg = false;
    if (s < end && (*s == '+' || *s == '-')) { neg = *s == '-'; ++s; }
    if (s == end) return false;
    int64_t v = 0;
    for (; s < end; ++s) {
        if (*s < '0' || *s > '9' || v > (INT64_MAX - 9) / 10) return false;
        v = v * 10 + (*s - '0');
    }
    *out = neg ? -v : v;
    return true;
}

// next whitespace-separated token of [p, end): returns the position after it, or null at the end of the line
inline const char *next_token(const char *p, const char *end, const char **tb, const char **te)
{
    while (p < end && is_space(*p)) ++p;
    if (p == end) return nullptr;
    *tb = p;
    while (p < end && !is_space(*p)) ++p;
    *te = p;
    return p;
}

int fail_line(int64_t line, const char *what)
{
    rt_set_error("OBJ line %lld: %s", (long long)line, what);
    return RT_ERR_INVALID;
}

int parse(const char *data, size_t size, ObjFile &f)
{
    int mesh = -1, material = -1;
    int64_t line_no = 0;
    const char *p = data, *const file_end = data + size;
    std::vector<int64_t> face; // (vi, ti, ni) of the current face's corners
    while (p < file_end) {
        const char *eol = p;
        while (eol < file_end && *eol != '\n' && *eol != '\r') ++eol;
        const char *line = p, *line_end = eol;
        p = eol < file_end ? eol + 1 : eol;
        ++line_no;
        if (line == line_end || *line == '#') continue;
        const char *kb, *ke;
        const char *q = next_token(line, line_end, &kb, &ke);
        if (!q) continue;
        const size_t klen = (size_t)(ke - kb);
        if (klen == 1 && *kb == 'v') {
            double v[3];
            for (int k = 0; k < 3; ++k) {
                const char *tb, *te;
                q = next_token(q, line_end, &tb, &te);
                if (!q || !parse_double(tb, te, &v[k])) return fail_line(line_no, "v needs three numbers");
            }
            f.pos.push_back((float)v[0]); f.pos.push_back((float)v[1]); f.pos.push_back((float)v[2]);
        } else if (klen == 2 && kb[0] == 'v' && kb[1] == 'n') {
            double v[3];
            for (int k = 0; k < 3; ++k) {
                const char *tb, *te;
                q = next_token(q, line_end, &tb, &te);
                if (!q || !parse_double(tb, te, &v[k])) return fail_line(line_no, "vn needs three numbers");
            }
            f.nrm.push_back((float)v[0]); f.nrm.push_back((float)v[1]); f.nrm.push_back((float)v[2]);
        } else if (klen == 2 && kb[0] == 'v' && kb[1] == 't') {
            double v[2] = {0.0, 0.0};
            const char *tb, *te;
            q = next_token(q, line_end, &tb, &te);
            if (!q || !parse_double(tb, te, &v[0])) return fail_line(line_no, "vt needs a number");
            q = next_token(q, line_end, &tb, &te);
            if (q && !parse_double(tb, te, &v[1])) return fail_line(line_no, "vt: bad second coordinate");
            f.tex.push_back((float)v[0]); f.tex.push_back((float)v[1]);
        } else if (klen == 1 && *kb == 'o') {
            f.meshes.emplace_back();
            mesh = (int)f.meshes.size() - 1;
        } else if (klen == 6 && memcmp(kb, "usemtl", 6) == 0) {
            const char *tb, *te;
            std::string name = next_token(q, line_end, &tb, &te) ? std::string(tb, te) : std::string("\0(none)", 7);
            auto it = f.by_name.find(name);
            if (it == f.by_name.end()) {
                f.materials.emplace_back();
                it = f.by_name.emplace(name, (int)f.materials.size() - 1).first;
            }
            material = it->second;
            if (mesh >= 0) {
                auto &ms = f.meshes[mesh].materials;
                bool have = false;
                for (int m : ms) have = have || m == material;
                if (!have) ms.push_back(material);
            }
        } else if (klen == 1 && *kb == 'f') {
            if (mesh < 0) { f.meshes.emplace_back(); mesh = (int)f.meshes.size() - 1; }
            if (material < 0) {
                auto it = f.by_name.find("default0");
                if (it == f.by_name.end()) {
                    f.materials.emplace_back();
                    it = f.by_name.emplace("default0", (int)f.materials.size() - 1).first;
                }
                material = it->second;
            }
            {
                auto &ms = f.meshes[mesh].materials;
                bool have = false;
                for (int m : ms) have = have || m == material;
                if (!have) ms.push_back(material);
            }
            face.clear();
            const int64_t np = (int64_t)f.pos.size() / 3, nt = (int64_t)f.tex.size() / 2, nn = (int64_t)f.nrm.size() / 3;
            const char *tb, *te;
            while ((q = next_token(q, line_end, &tb, &te)) != nullptr) {
                int64_t idx[3] = {-1, -1, -1};
                const char *s = tb;
                for (int k = 0; k < 3 && s <= te; ++k) {
                    const char *slash = s;
                    while (slash < te && *slash != '/') ++slash;
                    if (slash > s) {
                        int64_t raw;
                        if (!parse_int(s, slash, &raw)) return fail_line(line_no, "face corner is not v, v/t, v//n or v/t/n");
                        const int64_t count = k == 0 ? np : (k == 1 ? nt : nn);
                        idx[k] = raw > 0 ? raw - 1 : count + raw;
                    } else if (k == 0) {
                        return fail_line(line_no, "face corner without a vertex index");
                    }
                    if (slash == te) break;
                    s = slash + 1;
                }
                face.push_back(idx[0]); face.push_back(idx[1]); face.push_back(idx[2]);
            }
            const int64_t nc = (int64_t)face.size() / 3;
            ObjMaterial &mat = f.materials[material];
            if (nc == 0) {
                if (mat.format < 0) return fail_line(line_no, "face without corners");
                continue;
            }
            if (mat.format < 0) mat.format = (face[1] >= 0 ? 1 : 0) | (face[2] >= 0 ? 2 : 0);
            for (int64_t i = 2; i < nc; ++i) { // fan: (c0, c[i-1], c[i])
                const int64_t tri[3] = {0, i - 1, i};
                for (int64_t c : tri) { mat.corners.push_back(face[3 * c]); mat.corners.push_back(face[3 * c + 1]); mat.corners.push_back(face[3 * c + 2]); }
                f.meshes[mesh].n_faces += 1;
            }
        }
    }
    for (int m = 0; m < (int)f.meshes.size(); ++m)
        if (!f.meshes[m].materials.empty()) f.listed.push_back(m);
    return RT_OK;
}

ObjFile *from_handle(uint64_t h) { return reinterpret_cast<ObjFile *>((uintptr_t)h); }

} // namespace

extern "C" {

int rt_obj_load(const char *path, uint64_t *out_handle)
{
    RT_REQUIRE(path && out_handle, "path / handle");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { rt_set_error("cannot open %s: %s", path, strerror(errno)); return RT_ERR_INVALID; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); rt_set_error("cannot stat %s", path); return RT_ERR_INVALID; }
    ObjFile *f = new ObjFile();
    int rc = RT_OK;
    if (st.st_size > 0) {
        void *map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map == MAP_FAILED) { close(fd); delete f; rt_set_error("cannot map %s", path); return RT_ERR_INVALID; }
        madvise(map, (size_t)st.st_size, MADV_SEQUENTIAL);
        rc = parse((const char *)map, (size_t)st.st_size, *f);
        munmap(map, (size_t)st.st_size);
    }
    close(fd);
    if (rc != RT_OK) { delete f; return rc; }
    *out_handle = (uint64_t)(uintptr_t)f;
    return RT_OK;
}

int rt_obj_mesh_count(uint64_t handle)
{
    return handle ? (int)from_handle(handle)->listed.size() : 0;
}

int rt_obj_mesh_info(uint64_t handle, int mesh, int64_t *n_vertices, int64_t *n_faces, int *format)
{
    RT_REQUIRE(handle, "handle");
    ObjFile *f = from_handle(handle);
    RT_REQUIRE(mesh >= 0 && mesh < (int)f->listed.size(), "mesh index");
    const ObjMesh &m = f->meshes[f->listed[mesh]];
    const ObjMaterial &mat = f->materials[m.materials[0]];
    if (n_vertices) *n_vertices = (int64_t)mat.corners.size() / 3;
    if (n_faces) *n_faces = m.n_faces;
    if (format) *format = mat.format < 0 ? 0 : mat.format;
    return RT_OK;
}

int rt_obj_mesh_rows(uint64_t handle, int mesh, float *rows, int64_t row_floats)
{
    RT_REQUIRE(handle && rows, "handle / rows");
    RT_REQUIRE(row_floats >= 10, "a MeshVertex row has 20 floats (P@0 N@4 C@8)");
    ObjFile *f = from_handle(handle);
    RT_REQUIRE(mesh >= 0 && mesh < (int)f->listed.size(), "mesh index");
    const ObjMaterial &mat = f->materials[f->meshes[f->listed[mesh]].materials[0]];
    const int64_t n = (int64_t)mat.corners.size() / 3;
    const int64_t np = (int64_t)f->pos.size() / 3, nt = (int64_t)f->tex.size() / 2, nn = (int64_t)f->nrm.size() / 3;
    const bool has_t = mat.format >= 0 && (mat.format & 1), has_n = mat.format >= 0 && (mat.format & 2);
    // an absent index (-1) on a corner of a material whose format has that attribute reads the LAST table entry, like the
    // NumPy gather of the Python formulation; anything else out of range is an error
    for (int64_t i = 0; i < n; ++i) {
        int64_t vi = mat.corners[3 * i], ti = mat.corners[3 * i + 1], ni = mat.corners[3 * i + 2];
        float *r = rows + i * row_floats;
        if (vi < 0) vi += np;
        if (vi < 0 || vi >= np) { rt_set_error("OBJ: vertex index out of range"); return RT_ERR_INVALID; }
        r[0] = f->pos[3 * vi]; r[1] = f->pos[3 * vi + 1]; r[2] = f->pos[3 * vi + 2];
        if (has_n) {
            if (ni < 0) ni += nn;
            if (ni < 0 || ni >= nn) { rt_set_error("OBJ: normal index out of range"); return RT_ERR_INVALID; }
            r[4] = f->nrm[3 * ni]; r[5] = f->nrm[3 * ni + 1]; r[6] = f->nrm[3 * ni + 2];
        }
        if (has_t) {
            if (ti < 0) ti += nt;
            if (ti < 0 || ti >= nt) { rt_set_error("OBJ: texture coordinate index out of range"); return RT_ERR_INVALID; }
            r[8] = f->tex[2 * ti]; r[9] = f->tex[2 * ti + 1];
        }
    }
    return RT_OK;
}

int rt_obj_free(uint64_t handle)
{
    delete from_handle(handle);
    return RT_OK;
}

} // extern "C"
