// model.cpp - see model.h
#include <model.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <tuple>

#ifdef TRB_DEVICE_BACKEND
void trb_host_release_model(const Model& m);  // our_gl.cpp
#endif

Model::Model(const std::string& fn) : filename(fn) {
    size_t slash = fn.find_last_of("/\\");
    directory = slash == std::string::npos ? std::string(".") : fn.substr(0, slash);
}
Model::~Model() { unload(); }

Model::Model(const Model& o)
    : vertices(o.vertices), indices(o.indices), materials(o.materials), subMeshes(o.subMeshes),
      materialNames(o.materialNames), materialMaps(o.materialMaps), filename(o.filename), directory(o.directory),
      isLoaded(o.isLoaded), localAABB(o.localAABB) {}   // dev_* stay 0: the copy uploads its own buffers on first use
Model& Model::operator=(const Model& o) {
    if (this == &o) return *this;
    unload();                                            // releases this model's device buffers
    vertices = o.vertices; indices = o.indices; materials = o.materials;
    subMeshes = o.subMeshes; materialNames = o.materialNames; materialMaps = o.materialMaps;
    filename = o.filename; directory = o.directory; isLoaded = o.isLoaded; localAABB = o.localAABB;
    return *this;
}
Model::Model(Model&& o) noexcept
    : dev_mesh(o.dev_mesh), dev_diffuse(o.dev_diffuse), dev_normal(o.dev_normal), dev_specular(o.dev_specular),
      dev_uploaded(o.dev_uploaded), vertices(std::move(o.vertices)), indices(std::move(o.indices)),
      materials(std::move(o.materials)), subMeshes(std::move(o.subMeshes)), materialNames(std::move(o.materialNames)),
      materialMaps(std::move(o.materialMaps)), filename(std::move(o.filename)), directory(std::move(o.directory)),
      isLoaded(o.isLoaded), localAABB(o.localAABB) {
    o.dev_mesh = o.dev_diffuse = o.dev_normal = o.dev_specular = 0;
    o.dev_uploaded = false;
    o.isLoaded = false;
}
Model& Model::operator=(Model&& o) noexcept {
    if (this == &o) return *this;
    unload();
    vertices = std::move(o.vertices); indices = std::move(o.indices); materials = std::move(o.materials);
    subMeshes = std::move(o.subMeshes); materialNames = std::move(o.materialNames); materialMaps = std::move(o.materialMaps);
    filename = std::move(o.filename); directory = std::move(o.directory); isLoaded = o.isLoaded; localAABB = o.localAABB;
    dev_mesh = o.dev_mesh; dev_diffuse = o.dev_diffuse; dev_normal = o.dev_normal; dev_specular = o.dev_specular;
    dev_uploaded = o.dev_uploaded;
    o.dev_mesh = o.dev_diffuse = o.dev_normal = o.dev_specular = 0;
    o.dev_uploaded = false;
    o.isLoaded = false;
    return *this;
}

void Model::unload() {
#ifdef TRB_DEVICE_BACKEND
    if (dev_uploaded) trb_host_release_model(*this);
#endif
    vertices.clear();
    indices.clear();
    materials.clear();
    subMeshes.clear();
    materialNames.clear();
    materialMaps.clear();
    isLoaded = false;
}

void Model::computeAABB() {
    if (vertices.empty()) { localAABB = AABB(); return; }
    vec3 lo, hi;
    lo.x = lo.y = lo.z = 1e9;
    hi.x = hi.y = hi.z = -1e9;
    for (const Vertex& v : vertices) {
        lo.x = std::min(lo.x, v.position.x); lo.y = std::min(lo.y, v.position.y); lo.z = std::min(lo.z, v.position.z);
        hi.x = std::max(hi.x, v.position.x); hi.y = std::max(hi.y, v.position.y); hi.z = std::max(hi.z, v.position.z);
    }
    vec3 margin = (hi - lo) * 0.01;
    localAABB = AABB(lo - margin, hi + margin);
}

bool Model::load() {
    if (isLoaded) return true;
    if (!loadObj(filename)) {
        std::cerr << "Failed to load model: " << filename << std::endl;
        return false;
    }
    computeAABB();
    // materials: one per newmtl block (or a single default one, model.cpp:119-123); per texture slot the path the .mtl
    // names, else the fallback naming of model.cpp:252-262: <dir>/<stem>_diffuse.tga, _nm.tga, _spec.tga, _emission.tga
    materials.clear();
    materials.resize(std::max<size_t>(1, materialNames.size()));
    size_t slash = filename.find_last_of("/\\");
    std::string base = slash == std::string::npos ? filename : filename.substr(slash + 1);
    size_t dot = base.find_last_of('.');
    std::string stem = directory + "/" + (dot == std::string::npos ? base : base.substr(0, dot));
    auto try_load = [](TGAImage& img, const std::string& p) -> bool {
        std::ifstream probe(p, std::ios::binary);
        if (!probe.is_open()) return false;  // absent texture: the accessors fall back (model.cpp:416,429,447)
        probe.close();
        if (!img.read_tga_file(p)) { img = TGAImage(); return false; }
        return true;
    };
    for (size_t m = 0; m < materials.size(); ++m) {
        const MtlMaps maps = m < materialMaps.size() ? materialMaps[m] : MtlMaps();
        auto slot = [&](TGAImage& img, const std::string& named, const char* fallback) {
            if (!named.empty() && try_load(img, directory + "/" + named)) return;
            try_load(img, stem + fallback);
        };
        slot(materials[m].diffuse, maps.diffuse, "_diffuse.tga");
        slot(materials[m].normal, maps.normal, "_nm.tga");
        slot(materials[m].specular, maps.specular, "_spec.tga");
        slot(materials[m].emission, maps.emission, "_emission.tga");
    }
    isLoaded = true;
    return true;
}

bool Model::loadMtl(const std::string& path) {
    std::ifstream in(path);
    if (!in.is_open()) return false;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string tag, value;
        ss >> tag;
        if (tag.empty() || tag[0] == '#') continue;
        if (tag == "newmtl") {
            ss >> value;
            materialNames.push_back(value);
            materialMaps.emplace_back();
            continue;
        }
        if (materialMaps.empty()) continue;
        // the last token of the statement is the file name (options like -bm 1.0 come before it)
        std::string tok;
        while (ss >> tok) value = tok;
        MtlMaps& m = materialMaps.back();
        if (tag == "map_Kd") m.diffuse = value;
        else if (tag == "map_Bump" || tag == "map_bump" || tag == "bump" || tag == "norm" || tag == "map_Kn") m.normal = value;
        else if (tag == "map_Ks") m.specular = value;
        else if (tag == "map_Ke") m.emission = value;
    }
    return true;
}

bool Model::loadObj(const std::string& path) {
    std::ifstream in(path);
    if (!in.is_open()) return false;
    std::vector<vec3> P, N;
    std::vector<vec2> T;
    std::map<std::tuple<int, int, int>, unsigned int> weld;   // per sub-mesh
    std::string pending_name;          // set by o / g: the next faces open a new sub-mesh
    int pending_material = 0;
    bool open_new = true;
    std::string line;
    auto begin_submesh = [&]() {
        SubMesh sm;
        sm.name = pending_name;
        sm.startIndex = (unsigned int)indices.size();
        sm.materialIndex = pending_material;
        sm.vertexStart = (unsigned int)vertices.size();
        subMeshes.push_back(sm);
        weld.clear();                  // identical vertices are joined per mesh, not across meshes
        open_new = false;
    };
    while (std::getline(in, line)) {
        if (line.size() < 2) continue;
        std::istringstream ss(line);
        std::string tag;
        ss >> tag;
        if (tag == "v") {
            float x, y, z;  // Assimp stores floats (model.cpp:160-185): parse to float, widen later
            ss >> x >> y >> z;
            vec3 p; p.x = x; p.y = y; p.z = z;
            P.push_back(p);
        } else if (tag == "vn") {
            float x, y, z;
            ss >> x >> y >> z;
            vec3 n; n.x = x; n.y = y; n.z = z;
            N.push_back(n);
        } else if (tag == "vt") {
            float u = 0, v = 0;
            ss >> u >> v;
            vec2 t; t.x = u; t.y = (float)(1.0f - v);  // aiProcess_FlipUVs
            T.push_back(t);
        } else if (tag == "mtllib") {
            std::string name;
            ss >> name;
            loadMtl(directory + "/" + name);
        } else if (tag == "usemtl") {
            std::string name;
            ss >> name;
            int idx = 0;
            for (size_t m = 0; m < materialNames.size(); ++m)
                if (materialNames[m] == name) idx = (int)m;
            if (idx != pending_material || subMeshes.empty()) open_new = true;
            pending_material = idx;
        } else if (tag == "o" || tag == "g") {
            ss >> pending_name;
            open_new = true;
        } else if (tag == "f") {
            if (open_new) begin_submesh();
            SubMesh& sm = subMeshes.back();
            std::vector<unsigned int> poly;
            std::string tok;
            while (ss >> tok) {
                int vi = 0, ti = 0, ni = 0;
                const char* s = tok.c_str();
                vi = std::atoi(s);
                const char* s1 = std::strchr(s, '/');
                if (s1) {
                    if (s1[1] != '/') ti = std::atoi(s1 + 1);
                    const char* s2 = std::strchr(s1 + 1, '/');
                    if (s2) ni = std::atoi(s2 + 1);
                }
                auto fix = [](int i, size_t n) { return i > 0 ? i - 1 : (i < 0 ? (int)n + i : -1); };
                vi = fix(vi, P.size()); ti = fix(ti, T.size()); ni = fix(ni, N.size());
                if (vi < 0 || vi >= (int)P.size()) return false;
                auto key = std::make_tuple(vi, ti, ni);
                auto it = weld.find(key);
                if (it == weld.end()) {
                    Vertex v;
                    v.position = P[vi];
                    if (ti >= 0 && ti < (int)T.size()) { v.texcoord = T[ti]; sm.hasTexCoords = true; }
                    if (ni >= 0 && ni < (int)N.size()) { v.normal = N[ni]; sm.hasNormals = true; }
                    // the sub-mesh numbers its vertices from 0; flattening adds vertexStart (model.cpp:152, 195)
                    it = weld.emplace(key, (unsigned int)(vertices.size() - sm.vertexStart)).first;
                    vertices.push_back(v);
                }
                poly.push_back(sm.vertexStart + it->second);
            }
            for (size_t k = 1; k + 1 < poly.size(); ++k) {  // fan triangulation
                indices.push_back(poly[0]);
                indices.push_back(poly[k]);
                indices.push_back(poly[k + 1]);
            }
            sm.indexCount = (unsigned int)indices.size() - sm.startIndex;
        }
    }
    generateNormalsIfNeeded(false);
    return !vertices.empty() && !indices.empty();
}

void Model::generateNormalsIfNeeded(bool) {
    // model.cpp:269-312: as soon as ONE vertex has no usable normal, all normals are rebuilt from the faces
    bool needs = false;
    for (const Vertex& v : vertices)
        if (norm(v.normal) < 0.001) { needs = true; break; }
    if (!needs) return;
    for (Vertex& v : vertices) v.normal = vec3();
    for (size_t f = 0; f + 2 < indices.size(); f += 3) {
        Vertex &a = vertices[indices[f]], &b = vertices[indices[f + 1]], &c = vertices[indices[f + 2]];
        vec3 n = cross(b.position - a.position, c.position - a.position);  // area weighted
        a.normal = a.normal + n; b.normal = b.normal + n; c.normal = c.normal + n;
    }
    for (Vertex& v : vertices) {
        vec3 n = vec3();
        if (norm(v.normal) > 0.001) n = normalized(v.normal); else n.z = 1;   // model.cpp:305-311
        v.normal.x = (float)n.x; v.normal.y = (float)n.y; v.normal.z = (float)n.z;  // keep fp32-representable
    }
}

vec3 Model::vert(int i) const {
    if (i < 0 || i >= (int)vertices.size()) return vec3();
    return vertices[i].position;
}
vec3 Model::vert(int iface, int nthvert) const {
    int k = iface * 3 + nthvert;
    if (k < 0 || k >= (int)indices.size()) return vec3();
    return vertices[indices[k]].position;
}
vec3 Model::normal(int iface, int nthvert) const {
    int k = iface * 3 + nthvert;
    if (k < 0 || k >= (int)indices.size()) { vec3 n; n.z = 1; return n; }
    return vertices[indices[k]].normal;
}
vec2 Model::uv(int iface, int nthvert) const {
    int k = iface * 3 + nthvert;
    if (k < 0 || k >= (int)indices.size()) return vec2();
    return vertices[indices[k]].texcoord;
}
static void texel_xy(const TGAImage& t, const vec2& uv, int& x, int& y) {  // model.cpp:420-423
    x = std::clamp(int(uv.x * t.width()), 0, t.width() - 1);
    y = std::clamp(int(uv.y * t.height()), 0, t.height() - 1);
}
TGAColor Model::diffuse(const vec2& uv) const {
    if (materials.empty() || !materials[0].hasDiffuse()) return TGAColor(255, 255, 255, 255);
    int x, y;
    texel_xy(materials[0].diffuse, uv, x, y);
    return materials[0].diffuse.get(x, y);
}
vec3 Model::normal(const vec2& uv) const {
    if (materials.empty() || !materials[0].hasNormal()) { vec3 n; n.z = 1; return n; }
    int x, y;
    texel_xy(materials[0].normal, uv, x, y);
    TGAColor c = materials[0].normal.get(x, y);
    vec3 n;
    n.x = (double)c[2] / 255.0 * 2.0 - 1.0;
    n.y = (double)c[1] / 255.0 * 2.0 - 1.0;
    n.z = (double)c[0] / 255.0 * 2.0 - 1.0;
    return normalized(n);
}
float Model::specular(const vec2& uv) const {
    if (materials.empty() || !materials[0].hasSpecular()) return 1.0f;
    int x, y;
    texel_xy(materials[0].specular, uv, x, y);
    TGAColor c = materials[0].specular.get(x, y);
    return c[0] / 255.0f;
}
