// model_manager.h - ModelManager with the reference's interface (model_manager.h:9-39): a
// process-wide cache path -> weak_ptr<Model>, mutex guarded.  On the B200 backend it is what owns
// the lifetime of the device-resident vertex / index / texture buffers: they are uploaded once
// when a model is first drawn and released with the last shared_ptr.
#pragma once
#include <model.h>

#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>

class ModelManager {
public:
    static ModelManager& getInstance();
    std::shared_ptr<Model> loadModel(const std::string& path);  // nullptr + cerr on failure (model_manager.cpp:26-29)
    void unloadModel(const std::string& path);
    void cleanupUnused();
    size_t getLoadedCount() const;
    void printStats() const;

private:
    ModelManager() {}
    ModelManager(const ModelManager&) = delete;
    ModelManager& operator=(const ModelManager&) = delete;
    static std::string canonical(const std::string& path);
    std::unordered_map<std::string, std::weak_ptr<Model>> cache;
    mutable std::mutex mtx;
};
