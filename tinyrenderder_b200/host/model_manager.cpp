#include <model_manager.h>

#include <cstdlib>
#include <iostream>
#include <limits.h>

ModelManager& ModelManager::getInstance() {
    static ModelManager inst;
    return inst;
}
std::string ModelManager::canonical(const std::string& path) {
    char buf[PATH_MAX];
    if (realpath(path.c_str(), buf)) return std::string(buf);
    return path;
}
std::shared_ptr<Model> ModelManager::loadModel(const std::string& path) {
    std::lock_guard<std::mutex> lock(mtx);
    const std::string key = canonical(path);
    auto it = cache.find(key);
    if (it != cache.end())
        if (auto alive = it->second.lock()) return alive;
    auto model = std::make_shared<Model>(path);
    if (!model->load()) {
        std::cerr << "ModelManager: Failed to load model: " << path << std::endl;
        return nullptr;
    }
    cache[key] = model;
    return model;
}
void ModelManager::unloadModel(const std::string& path) {
    std::lock_guard<std::mutex> lock(mtx);
    cache.erase(canonical(path));
}
void ModelManager::cleanupUnused() {
    std::lock_guard<std::mutex> lock(mtx);
    for (auto it = cache.begin(); it != cache.end();)
        it = it->second.expired() ? cache.erase(it) : std::next(it);
}
size_t ModelManager::getLoadedCount() const {
    std::lock_guard<std::mutex> lock(mtx);
    size_t n = 0;
    for (auto& kv : cache) n += kv.second.expired() ? 0 : 1;
    return n;
}
void ModelManager::printStats() const {
    std::lock_guard<std::mutex> lock(mtx);
    std::cout << "=== ModelManager Stats ===" << std::endl;
    for (auto& kv : cache)
        if (auto m = kv.second.lock())
            std::cout << "  " << kv.first << ": " << m->nverts() << " vertices, " << m->nfaces() << " faces, refs "
                      << m.use_count() - 1 << std::endl;
}
