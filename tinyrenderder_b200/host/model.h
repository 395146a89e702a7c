// model.h - Model with the accessor set of the reference (model.h:62-88), loading Wavefront OBJ +
// <stem>_diffuse/_nm/_spec.tga without Assimp (absent here).  Written against the API
// subset shared by the reference's geometry.h/tgaimage.h and ours, so the same file compiles in
// the oracle build (reference headers + reference our_gl.cpp) and in the device build.
//
// Loader conventions (the reference delegates these to Assimp, whose output order is unpinned):
// faces in file order, polygons fan-triangulated, uv.y := 1 - uv.y (aiProcess_FlipUVs, model.cpp:93).
// A new SUB-MESH starts at every `usemtl` / `o` / `g` statement that is followed by faces (Assimp makes one aiMesh per
// object and material); inside a sub-mesh there is one vertex per distinct (v,vt,vn) tuple in first-seen order
// (aiProcess_JoinIdenticalVertices works per mesh), and the sub-meshes are flattened into ONE vertex / index array with
// the sub-mesh's vertexStart added to its indices (processMesh, model.cpp:143-205).  Materials are the `newmtl` blocks of
// the `mtllib` file in file order (a file without one gets a single default material); a material's textures come from
// its map_Kd / map_Bump (bump, norm) / map_Ks / map_Ke statements and, when the .mtl names none or the file does not load,
// from <stem>_diffuse.tga / _nm.tga / _spec.tga / _emission.tga next to the model (loadTexture, model.cpp:228-267).  Like
// the reference only materials[0] is ever sampled (model.cpp:416-459).  When any vertex lacks a normal, ALL normals are
// regenerated from the faces (generateNormalsIfNeeded, model.cpp:269-312).
#pragma once
#include <geometry.h>
#include <tgaimage.h>

#include <cstdint>
#include <string>
#include <vector>

struct Vertex {  // model.h:14-20 (tangent space is computed by the reference but never read)
    vec3 position, normal;
    vec2 texcoord;
    vec3 tangent, bitangent;
};
struct SubMesh {  // model.h:23-31: a run of the flattened index array that shares a material
    std::string name;
    unsigned int startIndex = 0, indexCount = 0;
    int materialIndex = 0;
    unsigned int vertexStart = 0;   // offset added to the sub-mesh's own vertex numbering (model.cpp:152, 195)
    bool hasNormals = false, hasTexCoords = false;
};
struct MaterialTextures {  // model.h:34-45
    TGAImage diffuse, normal, specular, emission;
    bool hasDiffuse() const { return diffuse.width() > 0; }
    bool hasNormal() const { return normal.width() > 0; }
    bool hasSpecular() const { return specular.width() > 0; }
    bool hasEmission() const { return emission.width() > 0; }
};

class Model {
public:
    explicit Model(const std::string& filename);
    ~Model();
    // The reference's Model is a plain copyable value type, so ported code may copy it.  The host arrays are
    // copied; the device handles are NOT shared (two owners would free them twice): a copy starts without
    // device buffers and uploads its own on first use, a move takes them over.
    Model(const Model& o);
    Model& operator=(const Model& o);
    Model(Model&& o) noexcept;
    Model& operator=(Model&& o) noexcept;
    bool load();
    void unload();

    int getVertexCount() const { return (int)vertices.size(); }
    int getIndexCount() const { return (int)indices.size(); }
    int getMaterialCount() const { return (int)materials.size(); }
    int getSubMeshCount() const { return (int)subMeshes.size(); }
    const SubMesh& getSubMesh(int i) const { return subMeshes[i]; }
    int nverts() const { return getVertexCount(); }
    int nfaces() const { return getIndexCount() / 3; }
    bool hasNormalMap() const { return !materials.empty() && materials[0].hasNormal(); }

    vec3 vert(int i) const;                       // model.cpp:391-394
    vec3 vert(int iface, int nthvert) const;      // model.cpp:396-400
    vec3 normal(int iface, int nthvert) const;    // model.cpp:402-406
    vec2 uv(int iface, int nthvert) const;        // model.cpp:408-412
    TGAColor diffuse(const vec2& uv) const;       // model.cpp:415-425
    vec3 normal(const vec2& uv) const;            // model.cpp:428-444
    float specular(const vec2& uv) const;         // model.cpp:446-459

    vec3 getCenter() const { return localAABB.getCenter(); }
    vec3 getSize() const { return localAABB.max - localAABB.min; }
    const AABB& getLocalAABB() const { return localAABB; }
    AABB getWorldAABB(const mat<4, 4>& m) const { return localAABB.transform(m); }
    const MaterialTextures& getMaterial(int i) const { return materials[i]; }
    const std::vector<Vertex>& getVertices() const { return vertices; }
    const std::vector<unsigned int>& getIndices() const { return indices; }
    const std::string& path() const { return filename; }

    // B200 backend: vertex / index / texture buffers live in HBM, uploaded once on first use
    // (0 = not uploaded).  Filled by our_gl.cpp; unused in the oracle build.
    mutable std::uint64_t dev_mesh = 0, dev_diffuse = 0, dev_normal = 0, dev_specular = 0;
    mutable bool dev_uploaded = false;

private:
    bool loadObj(const std::string& path);
    void generateNormalsIfNeeded(bool had_normals);
    void computeAABB();                            // model.cpp:15-40 (+1 % margin)
    std::vector<Vertex> vertices;
    std::vector<unsigned int> indices;
    std::vector<MaterialTextures> materials;
    std::vector<SubMesh> subMeshes;
    std::vector<std::string> materialNames;                       // order of the newmtl statements (material index)
    struct MtlMaps { std::string diffuse, normal, specular, emission; };
    std::vector<MtlMaps> materialMaps;                            // texture paths named by the .mtl file
    bool loadMtl(const std::string& path);
    std::string filename, directory;
    bool isLoaded = false;
    AABB localAABB;
};
