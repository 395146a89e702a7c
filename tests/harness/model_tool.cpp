// model_tool.cpp - load a model with the host layer's loader and print what it built (test tool).
// usage: model_tool file.obj
#include <model.h>
#include <cstdio>
int main(int argc, char** argv) {
    if (argc < 2) return 2;
    Model m(argv[1]);
    if (!m.load()) return 1;
    printf("vertices %d indices %d submeshes %d materials %d\n", m.getVertexCount(), m.getIndexCount(), m.getSubMeshCount(),
           m.getMaterialCount());
    for (int i = 0; i < m.getSubMeshCount(); ++i) {
        const SubMesh& s = m.getSubMesh(i);
        printf("submesh %d name %s start %u count %u material %d vertexStart %u normals %d uvs %d\n", i, s.name.c_str(), s.startIndex,
               s.indexCount, s.materialIndex, s.vertexStart, (int)s.hasNormals, (int)s.hasTexCoords);
    }
    for (int i = 0; i < m.getMaterialCount(); ++i) {
        const MaterialTextures& t = m.getMaterial(i);
        printf("material %d diffuse %d normal %d specular %d emission %d\n", i, t.diffuse.width(), t.normal.width(), t.specular.width(),
               t.emission.width());
    }
    printf("indices");
    for (unsigned int k : m.getIndices()) printf(" %u", k);
    printf("\n");
    for (const Vertex& v : m.getVertices())
        printf("v %.9g %.9g %.9g n %.9g %.9g %.9g t %.9g %.9g\n", v.position.x, v.position.y, v.position.z, v.normal.x, v.normal.y,
               v.normal.z, v.texcoord.x, v.texcoord.y);
    return 0;
}
