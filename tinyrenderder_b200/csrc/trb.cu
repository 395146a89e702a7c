// trb.cu - host side of the B200 rasterization backend: the C ABI of include/trb.h over the
// kernels of kernels.cuh.  One context = one GPU = one CUDA stream.  No CPU fallback: every
// entry point that needs the device fails with TRB_E_CUDA when it is not there.
#include "../../include/trb.h"
#include "kernels.cuh"
#include "tga_rle.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace trbk;

// mesh_order.cu: coherent processing order of a large indexed mesh (Morton order of the triangle centroids)
size_t trb_mesh_order_scratch_bytes(uint32_t ntris);
cudaError_t trb_mesh_order_build(const float4* pos4, uint32_t nverts, const uint32_t* idx, uint32_t ntris, uint32_t* perm_out,
                                 uint32_t* idx_perm_out, void* scratch, size_t scratch_bytes, int sms, cudaStream_t st);
size_t trb_vertex_order_scratch_bytes(uint32_t nverts);
cudaError_t trb_vertex_order_apply(const float4* pos_in, const float* attr_in, uint32_t nverts, uint32_t* idx, uint64_t nidx,
                                   float4* pos_out, float* attr_out, void* scratch, size_t scratch_bytes, int sms, cudaStream_t st);
cudaError_t trb_soup_order_apply(const float4* pos_in, const float* attr_in, uint32_t nverts, const uint32_t* perm, uint32_t ntris,
                                 float4* pos_out, float* attr_out, uint32_t* inv_perm, cudaStream_t st);

namespace {

// Frame recordings (trb_record_begin .. trb_replay, below) bake device addresses into a CUDA graph.  g_generation moves
// whenever an address a recording may hold stops being valid (a scratch buffer is reallocated or released, a mesh or
// texture is freed): a replay checks that it has not moved since the recording ended.  While a thread records, buffers
// may not grow (growth synchronises the stream, which a capturing stream cannot): the frame is rendered once first.
std::atomic<uint64_t> g_generation{0};
thread_local int t_recording = 0;      // contexts recording on this thread (the host layer records on several at once)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes, cudaStream_t st) {
        if (bytes <= cap) return cudaSuccess;
        if (t_recording) return cudaErrorNotPermitted;
        if (p) {
            ++g_generation;             // a first allocation invalidates nothing: no graph can hold an address that did not exist
            cudaStreamSynchronize(st);  // kernels in flight may still read the old block
            cudaFree(p);
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) {
            cudaFree(p);
            ++g_generation;
        }
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Frame arena: bump allocator for everything that must live until the next flush (vertex
// records, per-draw matrices and uniforms).  Reset at begin_frame / after flush; grows by chaining.
struct Arena {
    struct Block { char* p; size_t cap, used; };
    std::vector<Block> blocks;
    size_t next_cap = (size_t)64 << 20;
    void* alloc(size_t bytes, cudaError_t& err) {
        bytes = (bytes + 255) & ~(size_t)255;
        for (auto& b : blocks)
            if (b.cap - b.used >= bytes) {
                void* r = b.p + b.used;
                b.used += bytes;
                return r;
            }
        if (t_recording) {       // a recorded frame lives in the blocks the frame before it sized
            err = cudaErrorNotPermitted;
            return nullptr;
        }
        size_t cap = next_cap;
        while (cap < bytes) cap *= 2;
        Block nb{nullptr, cap, 0};
        err = cudaMalloc((void**)&nb.p, cap);
        if (err != cudaSuccess) return nullptr;
        next_cap = cap * 2;
        nb.used = bytes;
        blocks.push_back(nb);
        return nb.p;
    }
    void reset() {
        for (auto& b : blocks) b.used = 0;
    }
    void release() {
        for (auto& b : blocks) cudaFree(b.p);
        blocks.clear();
    }
};

// Resource blocks (meshes, textures) are recycled instead of cudaFree'd: cudaFree synchronises the
// whole device and a caller that re-uploads its scene every frame would pay it each time.  A block
// is tagged with an event recorded on the render stream when it is freed, and with the frame it
// was freed in; uploads run on their own stream, so a block is handed out again once that event has
// completed (the kernels that read it are done).  When no matching block is complete, one freed in an
// EARLIER frame is handed out with its event for the upload stream to wait on (the upload then starts
// when that older frame has rendered, i.e. it still overlaps the frame in flight); blocks freed in the
// current frame are left alone and a new block is allocated instead.  A caller that uploads, renders
// and frees every frame therefore settles into two alternating sets, and the upload of frame n+1
// overlaps the rendering of frame n.
struct BlockCache {
    struct Entry { size_t bytes; void* p; cudaEvent_t freed; uint64_t frame; };
    std::vector<Entry> free_blocks;
    std::vector<cudaEvent_t> ev_pool;
    uint64_t frame = 0;                        // advanced by begin_batch
    size_t cached_bytes = 0;
    static constexpr size_t SOFT_LIMIT = (size_t)16 << 30;   // beyond this, wait for a busy block rather than grow
    // *wait_for: event the consumer stream must wait on before writing the block (nullptr: none).
    // A new block comes from the stream-ordered allocator on the consumer stream: unlike cudaMalloc it
    // does not wait for the device, so growing the second set in the middle of a frame loop costs a
    // driver call, not a pipeline drain (measured: ~60 ms of stalls over the first steps with cudaMalloc).
    cudaError_t get(void** out, size_t bytes, cudaEvent_t* wait_for, cudaStream_t consumer) {
        bytes = (bytes + 511) & ~(size_t)511;
        *wait_for = nullptr;
        int busy = -1;
        for (size_t i = 0; i < free_blocks.size(); ++i) {
            Entry& e = free_blocks[i];
            if (e.bytes < bytes || e.bytes > bytes + bytes / 8 + 4096) continue;
            if (cudaEventQuery(e.freed) == cudaSuccess) {
                *out = e.p;
                take(i);
                return cudaSuccess;
            }
            if (busy < 0 && (e.frame < frame || cached_bytes > SOFT_LIMIT)) busy = (int)i;
        }
        (void)cudaGetLastError();   // cudaErrorNotReady from the queries is not an error
        if (busy >= 0) {            // oldest acceptable block: the consumer stream waits for its readers
            *out = free_blocks[busy].p;
            *wait_for = free_blocks[busy].freed;   // stays valid: returned to the pool, never destroyed before release()
            take((size_t)busy);
            return cudaSuccess;
        }
        if (cudaMallocAsync(out, bytes, consumer) == cudaSuccess) return cudaSuccess;
        (void)cudaGetLastError();
        return cudaMalloc(out, bytes);
    }
    void take(size_t i) {
        ev_pool.push_back(free_blocks[i].freed);
        cached_bytes -= free_blocks[i].bytes;
        free_blocks.erase(free_blocks.begin() + i);
    }
    void put(void* p, size_t bytes, cudaStream_t readers) {
        if (!p) return;
        cudaEvent_t ev = nullptr;
        if (ev_pool.size() > 8) {              // keep a few in reserve: an event handed out as wait_for may still be pending
            ev = ev_pool.front();
            ev_pool.erase(ev_pool.begin());
        } else {
            cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        }
        cudaEventRecord(ev, readers);
        bytes = (bytes + 511) & ~(size_t)511;
        free_blocks.push_back(Entry{bytes, p, ev, frame});
        cached_bytes += bytes;
    }
    void release() {
        for (auto& b : free_blocks) {
            cudaFree(b.p);
            cudaEventDestroy(b.freed);
        }
        for (auto e : ev_pool) cudaEventDestroy(e);
        free_blocks.clear();
        ev_pool.clear();
        cached_bytes = 0;
    }
};

// Pinned staging ring of the upload stream: the host fills a chunk (interleaving vertex attributes
// on the way) while earlier chunks are in flight; it only ever waits for the chunk it is about to reuse.
struct UploadRing {
    static constexpr size_t CHUNK = (size_t)8 << 20;
    static constexpr int SLOTS = 4;
    char* p[SLOTS] = {};
    cudaEvent_t done[SLOTS] = {};
    bool used[SLOTS] = {};
    int next = 0;
    cudaError_t slot(char** out, cudaEvent_t* ev) {
        const int i = next;
        next = (next + 1) % SLOTS;
        if (!p[i]) {
            cudaError_t e = cudaHostAlloc((void**)&p[i], CHUNK, cudaHostAllocDefault);
            if (e != cudaSuccess) return e;
            e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        } else if (used[i]) {
            cudaError_t e = cudaEventSynchronize(done[i]);
            if (e != cudaSuccess) return e;
        }
        used[i] = true;
        *out = p[i];
        *ev = done[i];
        return cudaSuccess;
    }
    void release() {
        for (int i = 0; i < SLOTS; ++i) {
            if (p[i]) {
                cudaEventSynchronize(done[i]);
                cudaFreeHost(p[i]);
                cudaEventDestroy(done[i]);
            }
            p[i] = nullptr;
            used[i] = false;
        }
    }
};

struct Mesh {
    float4* pos4 = nullptr;
    float* attr8 = nullptr;
    uint32_t* idx = nullptr;   // nullptr = implicit
    uint32_t nverts = 0;
    uint64_t nidx = 0;
    bool alive = false;
    // processing order of a large indexed mesh (mesh_order.cu): slot j of a draw holds triangle perm[j], whose vertex
    // indices are idx_perm[3j..3j+2]; nullptr for small meshes (their vertex records stay in L2 whatever the order)
    uint32_t* perm = nullptr;
    uint32_t* idx_perm = nullptr;
    // a large SOUP (idx == nullptr) has its vertex arrays themselves in processing order: slot j = vertices 3j .. 3j + 2
    // (idx_perm stays nullptr), inv_perm[t] = slot of triangle t for the shade pass.  Every draw of it runs over the slots.
    uint32_t* inv_perm = nullptr;
    // trb_draw_shard inside a composite group: which vertices this rank's share refers to (built at the first such draw,
    // kept for the (shard count, rank, block size) it was built for)
    uint8_t* vmark = nullptr;
    uint32_t vmark_n = 0, vmark_r = 0, vmark_shift = 0;
};
struct Tex {
    uint8_t* px = nullptr;
    int w = 0, h = 0, bpp = 0;
    bool alive = false;
};

struct ShadowMap {
    DevBuf keys;
    int w = 0, h = 0;
};

struct ProfEntry {
    const char* name;
    cudaEvent_t a, b;
};
struct ProfAcc {
    std::string name;
    uint64_t launches = 0;
    double ms = 0;
};

}  // namespace

// One recorded frame (trb_record_begin .. trb_record_end): the launch sequence as an instantiated CUDA graph, the pinned
// block its host-to-device parameter copies read from, where in that block each draw's matrices / uniforms sit (so that
// trb_replay can rewrite them for a new camera), and the context state the frame leaves behind.
struct Recording {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    char* params = nullptr;             // pinned; the graph's memcpy nodes read it when the graph RUNS
    size_t params_cap = 0, params_used = 0;
    struct Draw {
        int kind, nviews;
        size_t mats_off, litf_off, uni_off;   // (size_t)-1: the draw has no such block
        size_t uni_host_bytes;                // size of the caller's uniform block per view
        std::vector<double> mv;               // the caller's ModelViews and uniform blocks as last passed
        std::vector<char> uni;
    };
    size_t maps_before = 0;             // shadow maps the context held when the recording began ...
    std::vector<ShadowMap> maps_kept;   // ... and the ones trb_keep_depth_as_shadow_map made inside it
    std::vector<Draw> draws;
    cudaEvent_t done = nullptr;         // the last replay has finished reading `params`
    std::string failed;                 // first error of a call inside the recording
    uint64_t generation = 0;
    uint64_t launches = 0;              // kernel launches of one replay
    // what the context looks like after the frame
    FrameDev frame{};
    DevBuf zkey, zsnap;                 // roles of the two depth planes (trb_depth_restore swaps them)
    bool have_snapshot = false, snap_stale = false, snap_lazy = false, snap_all = false, foreign_ids = false;
    uint64_t next_id = 0, tris_submitted = 0;
    int shade_row0 = 0, shade_row1 = -1;
};

struct TrbCtx {
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    std::string err;

    std::vector<Mesh> meshes;
    std::vector<Tex> textures;
    BlockCache cache;
    cudaStream_t upload_stream = nullptr; // H2D of meshes and textures, ordered against the render stream by events
    cudaEvent_t upload_ev = nullptr;
    UploadRing ring;

    // frame
    FrameDev frame{};
    bool in_frame = false;
    uint8_t clear[3] = {0, 0, 0};
    DevBuf zkey, vis, color, stats, zsnap, zlocal;
    bool have_snapshot = false;
    bool snap_stale = false;        // restored by pointer swap: zsnap must be refreshed before the key plane changes
    // tile-granular snapshot (default; TRB_LAZY_SNAPSHOT=0: the whole-plane copy + pointer swap above): zsnap only receives
    // the tiles the draws after the snapshot change, snap_saved holds one byte per tile slot (kernels.cuh k_snap_save)
    bool lazy_snapshot = true;
    bool collect_by_tiles = true;   // flushes inside the window collect their pixel list from the saved tiles (TRB_COLLECT_BY_TILES=0: whole plane)
    bool snap_lazy = false;         // the current snapshot is a tile-granular one
    bool snap_all = false;          // ... and every tile has been saved (nothing left to do per draw)
    DevBuf snap_saved;
    uint64_t lazy_max_tris = 1u << 20;   // draws with more triangles x views keep the direct path and save every tile first
    std::vector<ShadowMap> shadow_maps;
    std::vector<DevBuf> shadow_pool;   // planes of released shadow maps, recycled by trb_keep_depth_as_shadow_map
    Arena arena;
    std::vector<DrawDev> draws;     // since the last flush
    DevBuf draw_table;
    uint64_t next_id = 0;
    uint64_t tris_submitted = 0;
    int shade_row0 = 0, shade_row1 = -1;

    // per-draw scratch (stream ordered reuse)
    DevBuf shade_list, tribox, trirec, offsets, bins, scan_sums, scan_total, scratch_a, scratch_b;
    // what a draw needs zeroed, in ONE block so that one memset does it: the DrawCtl of the draw in flight, the direct
    // path's per-view list lengths, the per-tile counts and fill cursors
    DevBuf drawzero;
    DrawCtl* ctl_p = nullptr;
    uint32_t *direct_n_p = nullptr, *counts_p = nullptr, *cursor_p = nullptr;
    DevBuf heavy_list;               // tile slots of the draw's long bins
    // long bins cut into slices for the split flavour of k_raster_warp (kernels.cuh SplitArgs)
    DevBuf split_items, split_keys, split_ids;
    uint32_t* split_done_p = nullptr;   // per tile slot, inside the zeroed block
    bool split_on = true;               // TRB_SPLIT=0: long bins go to the CTA-per-tile kernel
    uint32_t split_s = 256;             // triangles per slice (TRB_SPLIT_S)
    uint32_t split_cap_fixed = 0;       // TRB_SPLIT_CAP: fixed capacity of the item list (tests force the spill-over with it)
    uint32_t split_hint = 0;            // slices recent draws asked for (+ 25 %)
    DevBuf direct_list;              // direct path: triangles that may own a pixel, per view
    DevBuf rle_work, rle_src, rle_out; // device-side TGA RLE encoder (tga_rle.cuh)
    bool sync_draws = false;         // TRB_SYNC_DRAWS=1: size the bins exactly (one stream sync per draw)
    uint64_t bin_hint = 0;           // entries: 1.25 x the largest R seen so far
    uint32_t bin_cap_fixed = 0;      // TRB_BIN_CAP: fixed bin capacity in entries (tests force the overflow path with it)
    uint32_t* host_total = nullptr;  // pinned + mapped: the scan kernel stores the bin total straight into it
    uint32_t* host_total_dev = nullptr;


    // fused P2P composite: peers' planes opened through CUDA IPC
    PeerPlanes peers{};
    void* peer_open[2 * MAX_PEERS] = {};
    int peer_rank = -1;

    // sort-last composite group (trb_comm_*): in-process members or IPC peers, device-side frame counters
    struct Comm {
        int n = 0, rank = -1;
        bool in_process = false;              // members[] valid: contexts of this process (trb_comm_init)
        TrbCtx* members[MAX_PEERS] = {};
        CommFlags* my_flags = nullptr;        // device memory of this rank (own 2 MB block: exported through CUDA IPC)
        const CommFlags* peer_flags[MAX_PEERS] = {};
        void* opened[3 * MAX_PEERS] = {};     // IPC mappings to close
        unsigned long long seq = 0;           // frames composited so far
        cudaEvent_t ev_drawn = nullptr, ev_done = nullptr;   // in-process groups on one device synchronise by events
        int W = 0, H = 0;
        const void* key_at_init = nullptr;
    } comm;

    // pipelined readback
    cudaStream_t copy_stream = nullptr;
    // three staging areas: with two, the host can only run one step ahead of the copy engine (it blocks on the slot of
    // step s-2 before it may queue step s), and its ~1 ms of enqueue work after that wait shows up as bubbles on both the
    // render stream and the copy engine whenever rendering and the D2H copy take about equally long (config 3 on one GPU)
    static constexpr int RB_SLOTS = 3;
    DevBuf rb[RB_SLOTS];
    cudaEvent_t rb_ready[RB_SLOTS] = {}, rb_done[RB_SLOTS] = {};
    bool rb_inflight[RB_SLOTS] = {};
    int rb_idx = 0;

    // asynchronous TGA frame writer (trb_encode_tga_async): two jobs in flight
    struct TgaJob {
        DevBuf out;                          // packets of every view + the views' byte offsets
        uint32_t* offs_host = nullptr;       // pinned: nviews + 1 offsets
        size_t offs_cap = 0;
        cudaEvent_t offs_ready = nullptr, done = nullptr;
        std::vector<uint8_t*> outs;
        uint64_t* sizes = nullptr;
        uint64_t capacity = 0;
        int nviews = 0, W = 0, H = 0;
        bool pending_copy = false, inflight = false;
    } tga[2];
    int tga_idx = 0;

    // timing
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    bool profiling = false;
    std::vector<ProfEntry> prof_pending;
    cudaEvent_t trace_base = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<ProfAcc> prof_acc;
    uint64_t launches = 0;
    int big_ns = BIG_NS_DEFAULT, small_min = SMALL_MIN_DEFAULT, large_ns = LARGE_NS_DEFAULT;
    int direct_area = DIRECT_AREA_DEFAULT;
    int direct_by_pixel = 1;             // ordered soups: the direct path's warp vote looks at pixels, not tiles (TRB_DIRECT_BY_PIXEL=0: tiles)
    uint32_t warp_max = WARP_MAX_DEFAULT;
    int rw_blocks = RW_BLOCKS_DEFAULT;   // k_raster_warp instantiation (resident CTAs per SM the registers are sized for)
    bool shade_exact = false;   // true: all-fp64 lighting (exact.cuh); false: fp32 lighting (fastshade.cuh)
    // TMA descriptors of the current frame's depth-key / id planes (k_raster_warp stages its tiles with them)
    TileMaps tile_maps{};
    const void* maps_key = nullptr;     // what the descriptors were built for: rebuilt when any of it changes
    const void* maps_vis = nullptr;
    int maps_w = 0, maps_h = 0, maps_views = 0;
    bool maps_ok = false;
    bool use_tma = true;                // TRB_TMA=0: LDG / STG staging
    bool foreign_ids = false;   // trb_set_triangle_id_base was used: the id plane may hold winners other ranks rasterised
    // indexed meshes of at least this many triangles get a processing order at upload (TRB_MESH_ORDER_MIN_TRIS; 0 = never).
    // Default: from 2 M triangles up - below that the vertex records of a draw (32 B each) sit in the 126 MB L2 anyway.
    uint64_t order_min_tris = 2ull << 20;
    bool share_vertex = true;    // draw_shard in a composite group: vertex stage over the share's vertices only (TRB_SHARE_VERTEX=0: all)
    // ... from this many ranks up.  Measured on config 4: 8 ranks 0.595 -> 0.570 ms per frame (vertex stage 0.100 -> 0.062 ms,
    // composite + 0.003); 2 ranks 1.053 -> 1.086 (75 % of the vertex warps still have a marked lane, and half of the
    // pixels a rank shades then transform their three vertices themselves)
    int share_vertex_min_ranks = 4;
    bool order_vertices = true;  // ... and their vertices a Morton numbering (TRB_MESH_ORDER_VERTICES=0: keep the caller's)
    uint32_t shard_shift = 12;  // trb_draw_shard deals the processing order out in blocks of 2^shard_shift triangles (TRB_SHARD_SHIFT)

    // frame recordings (CUDA graphs): `rec` is the one being recorded, if any
    Recording* rec = nullptr;
    std::vector<Recording*> recordings;   // slot = handle - 1; nullptr = freed
    uint64_t launches_at_record = 0;
};

namespace {

int fail(TrbCtx* c, int code, const std::string& msg) {
    if (c) {
        c->err = msg;
        if (c->rec && c->rec->failed.empty()) c->rec->failed = msg;   // a call of the recording failed: record_end reports it
    }
    return code;
}
int refused(TrbCtx* c, const char* msg) {   // a call that may not run inside a recording: nothing was captured, nothing is lost
    if (c) c->err = msg;
    return TRB_E_ARG;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ == cudaErrorNotPermitted && t_recording)                                            \
            return fail(c, TRB_E_ARG, "record: " #call " would have to grow a buffer - render the frame once before recording it"); \
        if (e_ != cudaSuccess)                                                                     \
            return fail(c, e_ == cudaErrorMemoryAllocation ? TRB_E_NOMEM : TRB_E_CUDA,             \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                       \
    } while (0)

struct Launch {  // RAII around one kernel launch: counts it and, when profiling, brackets it with events
    TrbCtx* c;
    const char* name;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    Launch(TrbCtx* c_, const char* n, cudaStream_t s = nullptr, bool kernel = true) : c(c_), name(n), st(s ? s : c_->stream) {
        if (kernel) ++c->launches;   // copies are bracketed for the timeline but are not kernel launches
        if (c->profiling) {
            a = get_event();
            b = get_event();
            cudaEventRecord(a, st);
        }
    }
    ~Launch() {
        if (c->profiling) {
            cudaEventRecord(b, st);
            c->prof_pending.push_back(ProfEntry{name, a, b});
        }
    }
    cudaEvent_t get_event() {
        if (!c->ev_pool.empty()) {
            cudaEvent_t e = c->ev_pool.back();
            c->ev_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};

void prof_collect(TrbCtx* c) {
    if (c->prof_pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    if (c->upload_stream) cudaStreamSynchronize(c->upload_stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    // TRB_TRACE=<file>: per-launch timeline (start relative to the first traced launch, duration) for
    // hunting gaps between kernels; developer aid, off by default
    FILE* trace = nullptr;
    if (const char* path = getenv("TRB_TRACE")) trace = fopen(path, "a");
    for (auto& p : c->prof_pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, p.a, p.b);
        if (trace) {
            if (!c->trace_base) {
                c->trace_base = p.a;
                p.a = nullptr;            // kept for the lifetime of the context
            }
            float t0 = 0;
            if (p.a) cudaEventElapsedTime(&t0, c->trace_base, p.a);
            fprintf(trace, "%.4f %.4f %s\n", t0, ms, p.name);
        }
        ProfAcc* acc = nullptr;
        for (auto& a : c->prof_acc)
            if (a.name == p.name) acc = &a;
        if (!acc) {
            c->prof_acc.push_back(ProfAcc{p.name, 0, 0});
            acc = &c->prof_acc.back();
        }
        acc->launches++;
        acc->ms += ms;
        if (p.a) c->ev_pool.push_back(p.a);
        c->ev_pool.push_back(p.b);
    }
    if (trace) fclose(trace);
    c->prof_pending.clear();
}

// TRB_TRACE: host wall-clock of the entry points that should never wait for the device (developer aid)
struct HostSpan {
    const char* name;
    std::chrono::steady_clock::time_point t0;
    bool on;
    explicit HostSpan(const char* n) : name(n), on(getenv("TRB_TRACE") != nullptr) {
        if (on) t0 = std::chrono::steady_clock::now();
    }
    ~HostSpan() {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < 0.2) return;
        if (FILE* f = fopen((std::string(getenv("TRB_TRACE")) + ".host").c_str(), "a")) {
            fprintf(f, "%.3f %s\n", ms, name);
            fclose(f);
        }
    }
};

inline unsigned blocks_for(unsigned long long n) { return (unsigned)((n + TPB - 1) / TPB); }

int check_device(TrbCtx* c) {
    CU(cudaSetDevice(c->device));
    return TRB_OK;
}

void recording_destroy(Recording* r) {
    if (!r) return;
    if (r->exec) cudaGraphExecDestroy(r->exec);
    if (r->graph) cudaGraphDestroy(r->graph);
    if (r->done) cudaEventDestroy(r->done);
    if (r->params) cudaFreeHost(r->params);
    delete r;
}
// Host-to-device copy of a small parameter block (matrices, uniforms, the draw table) behind the context's stream.  The
// source is pageable and transient: outside a recording cudaMemcpyAsync stages it before it returns.  While recording the
// bytes go to the recording's pinned block and the captured copy reads them from there each time the graph runs;
// *slot (optional) receives their offset so that trb_replay can overwrite them.
int param_copy(TrbCtx* c, void* dst, const void* src, size_t bytes, size_t* slot = nullptr) {
    if (slot) *slot = (size_t)-1;
    if (!c->rec) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        return TRB_OK;
    }
    Recording& r = *c->rec;
    const size_t off = (r.params_used + 15) & ~(size_t)15;
    if (off + bytes > r.params_cap) return fail(c, TRB_E_NOMEM, "record: the frame's parameter blocks exceed the recording's 4 MB");
    memcpy(r.params + off, src, bytes);
    r.params_used = off + bytes;
    if (slot) *slot = off;
    CU(cudaMemcpyAsync(dst, r.params + off, bytes, cudaMemcpyHostToDevice, c->stream));
    return TRB_OK;
}
// entry points that synchronise, allocate or touch other streams cannot run inside a recording
#define NOT_WHILE_RECORDING(c, what)                                                                       \
    do {                                                                                                   \
        if ((c) && (c)->rec) return refused(c, what ": not inside trb_record_begin / trb_record_end"); \
    } while (0)

// between a tile-granular snapshot and the next one (or the end of the frame), as long as single tiles are being saved
inline bool snapshot_window(const TrbCtx* c) { return c->have_snapshot && c->snap_lazy && !c->snap_all; }

// upload the draw table and run the shade kernel over [row0,row1)
int do_flush(TrbCtx* c) {
    if (!c->in_frame) return fail(c, TRB_E_ARG, "flush: no frame");
    if (c->draws.empty()) return TRB_OK;
    int rc = check_device(c);
    if (rc) return rc;
    size_t bytes = c->draws.size() * sizeof(DrawDev);
    CU(c->draw_table.ensure(bytes, c->stream));
    rc = param_copy(c, c->draw_table.p, c->draws.data(), bytes);
    if (rc) return rc;
    int r0 = 0, r1 = c->frame.H;
    if (c->shade_row1 >= 0) {
        r0 = c->shade_row0;
        r1 = c->shade_row1;
    }
    if (r1 > r0) {
        const FrameDev& f = c->frame;
        const unsigned long long n = (unsigned long long)(r1 - r0) * f.W;
        bool config2 = false;
        for (const DrawDev& d : c->draws) config2 |= d.kind >= 4;
        CU(c->shade_list.ensure((size_t)f.npix * f.nviews * 4, c->stream));
        {
            Launch L(c, "k_shade_decide");
            k_shade_decide<<<(f.nviews + TPB - 1) / TPB, TPB, 0, c->stream>>>(f, n);
        }
        if (snapshot_window(c) && c->collect_by_tiles) {
            // sparse views, inside a tile-granular snapshot window: what was drawn since the snapshot (which flushed) lies
            // in the tiles marked as saved - collect from those, not from the whole id plane
            const uint32_t nslots = (uint32_t)((size_t)f.nviews * f.ntiles);
            Launch L(c, "k_shade_collect_tiles");
            k_shade_collect_tiles<<<(nslots + TPB - 1) / TPB, TPB, 0, c->stream>>>(f, nslots, c->snap_saved.as<uint8_t>(), r0, r1,
                                                                                 c->shade_list.as<uint32_t>());
        } else {   // sparse views only (device-side predicate)
            const unsigned need = blocks_for((n + SHADE_PX_PER_THREAD - 1) / SHADE_PX_PER_THREAD);
            dim3 grid(std::min(need, std::max(1u, (unsigned)c->sms * 16 / (unsigned)f.nviews)), f.nviews);
            Launch L(c, "k_shade_collect");
            k_shade_collect<<<grid, TPB, 0, c->stream>>>(f, r0, r1, c->shade_list.as<uint32_t>());
        }
        const DrawDev* table = c->draw_table.as<DrawDev>();
        const int nd = (int)c->draws.size();
        const uint32_t* list = c->shade_list.as<uint32_t>();
        // fp32 lighting needs the LitF block of a mesh draw: immediate-mode lit triangles (rasterize() with
        // host-computed varyings) send the whole flush down the fp64 kernels
        bool fast = !c->shade_exact;
        for (const DrawDev& d : c->draws) fast &= !((d.kind == 1 || d.kind == 2 || d.kind == 4) && !d.litf);
        const int variant = (config2 ? 2 : 0) | (fast ? 1 : 0);   // template <C2, FAST>
        {
            unsigned per_view = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>(
                blocks_for(n), ((unsigned long long)c->sms * 3 * 4 + f.nviews - 1) / f.nviews));
            const dim3 grid(per_view, f.nviews);
            Launch L(c, "k_shade");
            switch (variant) {
                case 0: k_shade<false, false><<<grid, TPB, 0, c->stream>>>(f, table, nd, list); break;
                case 1: k_shade<false, true><<<grid, TPB, 0, c->stream>>>(f, table, nd, list); break;
                case 2: k_shade<true, false><<<grid, TPB, 0, c->stream>>>(f, table, nd, list); break;
                default: k_shade<true, true><<<grid, TPB, 0, c->stream>>>(f, table, nd, list); break;
            }
        }
        {   // dense views only
            const dim3 grid((unsigned)((n + (unsigned long long)TPB * SHADE_DENSE_PX - 1) / ((unsigned long long)TPB * SHADE_DENSE_PX)), f.nviews);
            Launch L(c, "k_shade_dense");
            switch (variant) {
                case 0: k_shade_dense<false, false><<<grid, TPB, 0, c->stream>>>(f, table, nd, r0, r1); break;
                case 1: k_shade_dense<false, true><<<grid, TPB, 0, c->stream>>>(f, table, nd, r0, r1); break;
                case 2: k_shade_dense<true, false><<<grid, TPB, 0, c->stream>>>(f, table, nd, r0, r1); break;
                default: k_shade_dense<true, true><<<grid, TPB, 0, c->stream>>>(f, table, nd, r0, r1); break;
            }
        }
    }
    CU(cudaGetLastError());
    // the pageable->device copy of the table is staged before cudaMemcpyAsync returns, so the
    // host vector may be cleared now; device memory of the arena is recycled at begin_frame only
    c->draws.clear();
    return TRB_OK;
}

int exclusive_scan(TrbCtx* c, const uint32_t* in, uint32_t n, uint32_t* out, uint32_t bin_capacity, const SplitArgs& sp) {
    uint32_t nblocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    CU(c->scan_sums.ensure((size_t)nblocks * 4, c->stream));
    {
        Launch L(c, "k_scan_partial");
        k_scan_partial<<<nblocks, TPB, 0, c->stream>>>(in, n, c->scan_sums.as<uint32_t>(), c->ctl_p, c->warp_max,
                                                       c->heavy_list.as<uint32_t>(), sp);
    }
    {
        Launch L(c, "k_scan_final");
        k_scan_final<<<nblocks, TPB, 0, c->stream>>>(in, n, c->scan_sums.as<uint32_t>(), out, c->host_total_dev,
                                                     c->ctl_p, bin_capacity);
    }
    CU(cudaGetLastError());
    return TRB_OK;
}

// after a pointer-swap restore the spare plane no longer holds the snapshot: copy it back before the
// key plane changes again
int refresh_snapshot(TrbCtx* c) {
    if (!c->have_snapshot || !c->snap_stale) return TRB_OK;
    size_t bytes = (size_t)c->frame.npix * c->frame.nviews * 8;
    CU(cudaMemcpyAsync(c->zsnap.p, c->zkey.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
    c->snap_stale = false;
    return TRB_OK;
}

// tile-granular snapshot: save the tiles a draw is about to change (`counts` / `ctl` of that draw), or every tile not
// saved yet (`counts == nullptr`: a writer that does not go through the bins - a draw too large to give up the direct
// path for, a composite)
int snapshot_save_tiles(TrbCtx* c, const uint32_t* counts, const DrawCtl* ctl) {
    const FrameDev& f = c->frame;
    const uint32_t nslots = (uint32_t)((size_t)f.nviews * f.ntiles);
    Launch L(c, "k_snap_save");
    k_snap_save<<<(nslots + TPB - 1) / TPB, TPB, 0, c->stream>>>(f, nslots, counts, ctl, counts ? 0 : 1, c->snap_saved.as<uint8_t>(),
                                                               c->zsnap.as<unsigned long long>());
    CU(cudaGetLastError());
    if (!counts) c->snap_all = true;
    return TRB_OK;
}

// Tensor maps of the frame's planes for the TMA tile staging of k_raster_warp.  cuTensorMapEncodeTiled is a driver
// entry point; it is looked up at run time so that the library carries no link-time dependency on libcuda.
bool tile_maps_for(TrbCtx* c) {
    const FrameDev& f = c->frame;
    if (!c->use_tma || (f.W & 3)) return false;            // row pitch of the id plane must be a multiple of 16 bytes
    if (c->maps_key == f.zkey && c->maps_vis == f.vis && c->maps_w == f.W && c->maps_h == f.H && c->maps_views == f.nviews)
        return c->maps_ok;
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = nullptr;
    static bool looked_up = false;
    if (!looked_up) {
        looked_up = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeTiled>(fn);
        (void)cudaGetLastError();
    }
    c->maps_key = f.zkey; c->maps_vis = f.vis; c->maps_w = f.W; c->maps_h = f.H; c->maps_views = f.nviews;
    c->maps_ok = false;
    if (!encode) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)f.W, (cuuint64_t)f.H, (cuuint64_t)f.nviews};
    const cuuint32_t box[3] = {TILE, TILE, 1}, estr[3] = {1, 1, 1};
    const cuuint64_t kstr[2] = {(cuuint64_t)f.W * 8, (cuuint64_t)f.npix * 8}, vstr[2] = {(cuuint64_t)f.W * 4, (cuuint64_t)f.npix * 4};
    const CUresult a = encode(&c->tile_maps.key, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*)f.zkey, dims, kstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const CUresult b = encode(&c->tile_maps.vis, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)f.vis, dims, vstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    c->maps_ok = a == CUDA_SUCCESS && b == CUDA_SUCCESS;
    return c->maps_ok;
}

// bin + rasterise one draw whose vertex records are already in `vrec`.
//
// Default (asynchronous): nothing here waits for the device.  The bin buffer is sized from an
// estimate (4 entries per triangle and view, grown from the R of earlier draws, which the scan
// kernel leaves in mapped host memory); the kernels after the scan take their work from the
// device-side DrawCtl, and a draw whose R does not fit is rasterised by the unbinned kernels.
// TRB_SYNC_DRAWS=1 restores the exact sizing: one stream synchronisation per draw to read R.
int raster_draw(TrbCtx* c, const GeomArgs& g) {
    const FrameDev& f = c->frame;
    if (g.ntris == 0) return TRB_OK;
    int rs = refresh_snapshot(c);
    if (rs) return rs;
    const size_t nslots = (size_t)f.nviews * f.ntiles;
    if (nslots >= 0xFFFFFFFFull) return fail(c, TRB_E_ARG, "draw: too many tiles x views");
    // Between a tile-granular snapshot and its restore a draw saves the tiles it is about to change.  The bins say which
    // ones, so such a draw does without the direct path (whose atomics run inside the set-up kernel, before anything is
    // known about the draw) - the same exact (depth, id) minimum either way.  A draw too large for that saves every tile.
    bool save_tiles = snapshot_window(c);
    if (save_tiles && (uint64_t)g.nslots * (uint64_t)f.nviews > c->lazy_max_tris) {
        if ((rs = snapshot_save_tiles(c, nullptr, nullptr))) return rs;
        save_tiles = false;
    }
    const int direct_area = save_tiles ? 0 : c->direct_area;
    const uint32_t ndslots = g.nslots;     // == g.ntris unless the mesh carries a processing order (GeomArgs::perm)
    CU(c->tribox.ensure((size_t)f.nviews * ndslots * sizeof(uint2), c->stream));
    CU(c->trirec.ensure((size_t)f.nviews * ndslots * sizeof(TriRec), c->stream));
    CU(c->offsets.ensure(nslots * 4, c->stream));
    CU(c->heavy_list.ensure(nslots * 4, c->stream));
    {
        const size_t o_dn = 256, o_counts = o_dn + (((size_t)f.nviews * 4 + 255) & ~(size_t)255),
                     o_cursor = o_counts + ((nslots * 4 + 255) & ~(size_t)255),
                     o_done = o_cursor + ((nslots * 4 + 255) & ~(size_t)255), bytes = o_done + nslots * 4;
        static_assert(sizeof(DrawCtl) <= 256, "DrawCtl sits in the first 256 bytes of the zeroed block");
        CU(c->drawzero.ensure(bytes, c->stream));
        char* z = c->drawzero.as<char>();
        c->ctl_p = reinterpret_cast<DrawCtl*>(z);
        c->direct_n_p = reinterpret_cast<uint32_t*>(z + o_dn);
        c->counts_p = reinterpret_cast<uint32_t*>(z + o_counts);
        c->cursor_p = reinterpret_cast<uint32_t*>(z + o_cursor);
        c->split_done_p = reinterpret_cast<uint32_t*>(z + o_done);
        CU(cudaMemsetAsync(z, 0, bytes, c->stream));
    }
    if (direct_area > 0) CU(c->direct_list.ensure((size_t)f.nviews * ndslots * 4, c->stream));
    uint32_t capacity = 0xFFFFFFFFu;   // synchronous draws size the buffer after the scan
    if (!c->sync_draws) {
        // R of the most recent draw the device has finished scanning: a hint, never waited for
        c->bin_hint = std::max<uint64_t>(c->bin_hint, (uint64_t)c->host_total[0] + c->host_total[0] / 4);
        uint64_t want = std::max<uint64_t>(c->bin_hint, 4ull * g.ntris * f.nviews + 4ull * nslots + 65536);
        want = std::min<uint64_t>(want, 0xFFFFFFF0ull);
        if (c->bin_cap_fixed) want = c->bin_cap_fixed;
        CU(c->bins.ensure((size_t)want * 4, c->stream));
        capacity = c->bin_cap_fixed ? c->bin_cap_fixed : (uint32_t)std::min<uint64_t>(c->bins.cap / 4, 0xFFFFFFF0ull);
    }
    dim3 tgrid(blocks_for(ndslots), f.nviews);
    {
        Launch L(c, "k_setup_count");
        const dim3 sgrid((ndslots + TPB * SETUP_CHUNKS - 1) / (TPB * SETUP_CHUNKS), f.nviews);
        k_setup_count<<<sgrid, TPB, 0, c->stream>>>(f, g, c->tribox.as<uint2>(), c->trirec.as<TriRec>(),
                                                     c->counts_p, direct_area, c->direct_list.as<uint32_t>(),
                                                     c->direct_n_p, (c->direct_by_pixel && g.perm && !g.idx_perm) ? 1 : 0);
    }
    if (direct_area > 0) {
        // full grid once a draw of this context has had direct candidates (mapped flag, never waited for), else a few CTAs
        Launch L(c, "k_direct_resolve");
        if (c->host_total[9]) {
            k_direct_resolve<false><<<tgrid, TPB, 0, c->stream>>>(f, g, c->direct_list.as<uint32_t>(), c->direct_n_p, c->host_total_dev + 9);
        } else {
            const dim3 dgrid(std::min<unsigned>(tgrid.x, std::max(1u, (unsigned)c->sms * 4 / (unsigned)f.nviews)), f.nviews);
            k_direct_resolve<true><<<dgrid, TPB, 0, c->stream>>>(f, g, c->direct_list.as<uint32_t>(), c->direct_n_p, c->host_total_dev + 9);
        }
    }
    CU(cudaGetLastError());
    // item list of the split warp kernel: sized from what recent draws asked for (never waited for); a draw that needs
    // more sends the bins that do not fit to the CTA-per-tile kernel
    // (the kernel is only launched once a draw of this context has had long bins: config 3 never pays for it)
    SplitArgs sp{};
    sp.split_s = c->split_s;
    if (c->split_on && c->warp_max > 0 && c->host_total[3])
        c->split_hint = std::max<uint32_t>(c->split_hint, (uint32_t)std::min<uint64_t>((uint64_t)c->host_total[3] + c->host_total[3] / 4 + 256, 1u << 20));
    if (c->split_on && c->warp_max > 0 && (c->split_hint || c->split_cap_fixed)) {
        uint32_t cap = std::min<uint32_t>(std::max<uint32_t>(c->split_hint, 4096u), 65536u);
        if (c->split_cap_fixed) cap = c->split_cap_fixed;
        cap = (cap + RW_WARPS - 1) / RW_WARPS * RW_WARPS;
        CU(c->split_items.ensure((size_t)cap * sizeof(uint2), c->stream));
        CU(c->split_keys.ensure((size_t)cap * TILE * TILE * 8, c->stream));
        CU(c->split_ids.ensure((size_t)cap * TILE * TILE * 4, c->stream));
        sp.items = c->split_items.as<uint2>();
        sp.done = c->split_done_p;
        sp.keys = c->split_keys.as<unsigned long long>();
        sp.ids = c->split_ids.as<uint32_t>();
        sp.cap = cap;
    }
    int rc = exclusive_scan(c, c->counts_p, (uint32_t)nslots, c->offsets.as<uint32_t>(), capacity, sp);
    if (rc) return rc;
    bool long_bins = true;             // unknown without a round trip: k_raster's persistent grid finds an empty list
    if (c->sync_draws) {
        // the total is stored by the scan kernel directly into mapped host memory: a cudaMemcpy would
        // queue on the device->host copy engine behind a pipelined readback (trb_readback_async)
        CU(cudaStreamSynchronize(c->stream));
        const uint32_t R = c->host_total[0];
        if (R == 0) return TRB_OK;
        long_bins = c->host_total[1] > c->warp_max || c->host_total[2];   // k_raster also carries the overflow fallback's depth pass
        if (!c->host_total[2]) CU(c->bins.ensure((size_t)R * 4, c->stream));   // overflow (R >= 2^32): the unbinned kernels draw it
    }
    if (save_tiles && (rc = snapshot_save_tiles(c, c->counts_p, c->ctl_p))) return rc;
    {
        Launch L(c, "k_fill");
        k_fill<<<tgrid, TPB, 0, c->stream>>>(f, ndslots, c->tribox.as<uint2>(), c->offsets.as<uint32_t>(),
                                            c->cursor_p, c->bins.as<uint32_t>(), c->ctl_p);
    }
    RasterArgs ra;
    ra.ntris = ndslots;
    // slot_gid adds perm[position of the slot] (a mesh triangle) or the slot itself
    ra.ids = SlotIds{g.perm, g.perm ? g.id_base - g.first_tri : g.id_base, g.shard_n, g.shard_r, g.shard_shift};
    ra.trirec = c->trirec.as<TriRec>();
    ra.tribox = c->tribox.as<uint2>();
    ra.counts = c->counts_p;
    ra.offsets = c->offsets.as<uint32_t>();
    ra.bins = c->bins.as<uint32_t>();
    ra.big_ns = c->big_ns;
    ra.small_min = c->small_min;
    ra.large_ns = c->large_ns;
    ra.warp_max = c->warp_max;
    ra.ctl = c->ctl_p;
    ra.heavy_list = c->heavy_list.as<uint32_t>();
    if (c->warp_max > 0) {   // bins of 1..warp_max triangles: one warp per tile
        Launch L(c, "k_raster_warp");
        const dim3 grid((f.ntiles + RW_WARPS - 1) / RW_WARPS, f.nviews);
        const int tma = tile_maps_for(c) ? 1 : 0;
        switch (c->rw_blocks) {
            case 8: k_raster_warp<8, false><<<grid, RW_WARPS * 32, 0, c->stream>>>(f, ra, c->tile_maps, tma, sp); break;
            case 7: k_raster_warp<7, false><<<grid, RW_WARPS * 32, 0, c->stream>>>(f, ra, c->tile_maps, tma, sp); break;
            default: k_raster_warp<6, false><<<grid, RW_WARPS * 32, 0, c->stream>>>(f, ra, c->tile_maps, tma, sp); break;
        }
    }
    if (sp.cap) {            // slices of the longer bins: one warp each, the last one of a tile folds them into the frame
        Launch L(c, "k_raster_split");
        k_raster_warp<6, true><<<sp.cap / RW_WARPS, RW_WARPS * 32, 0, c->stream>>>(f, ra, c->tile_maps, 0, sp);
    }
    if (long_bins) {         // longer bins: one CTA per tile, persistent grid over the device-side list
        // asynchronous draws: the full persistent grid, because it may have to carry the unbinned depth pass
        const unsigned grid = c->sync_draws && !c->host_total[2] ? (unsigned)std::min<size_t>(nslots, (size_t)c->sms * TRB_RASTER_MIN_BLOCKS)
                                                                 : (unsigned)c->sms * TRB_RASTER_MIN_BLOCKS;
        Launch L(c, "k_raster");
        k_raster<<<grid, TPB, 0, c->stream>>>(f, ra);
    }
    {   // stand-in for a draw that overflowed its bins or 32-bit offsets (exits at once otherwise): the id pass of the
        // unbinned fallback; its depth pass ran inside k_raster's grid
        const dim3 grid((unsigned)std::min<unsigned>(blocks_for((unsigned long long)ndslots * 32),
                                                    std::max(1u, (unsigned)c->sms * 8 / (unsigned)f.nviews)), f.nviews);
        Launch L(c, "k_unbinned_ids");
        k_unbinned<true><<<grid, TPB, 0, c->stream>>>(f, ndslots, ra.ids, c->tribox.as<uint2>(), c->trirec.as<TriRec>(), c->ctl_p);
    }
    CU(cudaGetLastError());
    return TRB_OK;
}

// the per-view uniform blocks of a lit draw as the device reads them (LitUniforms, or ShadowUniformsDev for SHADOW_PHONG);
// `out` stays empty for the shaders without uniforms
int build_uniforms(TrbCtx* c, int kind, const void* uniforms, size_t ubytes, int nviews, std::vector<char>& out) {
    out.clear();
    if (kind == TRB_SHADER_FLAT_BARY || kind == TRB_SHADER_DEPTH) return TRB_OK;
    const bool shadow = kind == TRB_SHADER_SHADOW_PHONG;
    if (kind != TRB_SHADER_PHONG && kind != TRB_SHADER_EYE && kind != TRB_SHADER_GOURAUD && !shadow)
        return fail(c, TRB_E_SHADER, "shader kind has no device implementation (no CPU fallback)");
    const size_t want = shadow ? sizeof(TrbShadowUniforms) : sizeof(TrbPhongUniforms);
    if (!uniforms || ubytes != want) return fail(c, TRB_E_ARG, "draw: uniform block size");
    auto tex = [&](TrbTex t, TexView& out) -> bool {
        out = TexView{nullptr, 0, 0, 0};
        if (t == 0) return true;
        if (t > c->textures.size() || !c->textures[t - 1].alive) return false;
        const Tex& x = c->textures[t - 1];
        out = TexView{x.px, x.w, x.h, x.bpp};
        return true;
    };
    auto lit_of = [&](const TrbPhongUniforms& u, LitUniforms& L) -> bool {
        L.key = D3{u.key_dir_eye[0], u.key_dir_eye[1], u.key_dir_eye[2]};
        L.fill = D3{u.fill_dir_eye[0], u.fill_dir_eye[1], u.fill_dir_eye[2]};
        L.rim = D3{u.rim_dir_eye[0], u.rim_dir_eye[1], u.rim_dir_eye[2]};
        L.normal_map_strength = u.normal_map_strength;
        return tex(u.diffuse, L.diffuse) && tex(u.normal, L.normal) && tex(u.specular, L.specular);
    };
    if (!shadow) {
        out.resize(sizeof(LitUniforms) * nviews);
        LitUniforms* host = reinterpret_cast<LitUniforms*>(out.data());
        const TrbPhongUniforms* u = (const TrbPhongUniforms*)uniforms;
        for (int v = 0; v < nviews; ++v)
            if (!lit_of(u[v], host[v])) return fail(c, TRB_E_ARG, "draw: bad texture handle");
        return TRB_OK;
    }
    out.resize(sizeof(ShadowUniformsDev) * nviews);
    ShadowUniformsDev* host = reinterpret_cast<ShadowUniformsDev*>(out.data());
    const TrbShadowUniforms* u = (const TrbShadowUniforms*)uniforms;
    for (int v = 0; v < nviews; ++v) {
        if (!lit_of(u[v].phong, host[v].lit)) return fail(c, TRB_E_ARG, "draw: bad texture handle");
        if (u[v].shadow_map < 0 || (size_t)u[v].shadow_map >= c->shadow_maps.size())
            return fail(c, TRB_E_ARG, "draw: bad shadow map index");
        const ShadowMap& sm = c->shadow_maps[u[v].shadow_map];
        if (u[v].shadow_w != sm.w || u[v].shadow_h != sm.h) return fail(c, TRB_E_ARG, "draw: shadow map size mismatch");
        ShadowParams& S = host[v].shadow;
        memcpy(S.lmv, u[v].light_modelview, 128);
        memcpy(S.lpr, u[v].light_perspective, 128);
        memcpy(S.lvp, u[v].light_viewport, 128);
        S.bias = u[v].shadow_bias;
        S.darkening = u[v].shadow_darkening;
        S.map_keys = sm.keys.as<unsigned long long>();
        S.map_z = nullptr;
        S.w = sm.w;
        S.h = sm.h;
    }
    return TRB_OK;
}
int resolve_uniforms(TrbCtx* c, int kind, const void* uniforms, size_t ubytes, int nviews, const void** dev_out,
                     size_t* slot = nullptr) {
    *dev_out = nullptr;
    if (slot) *slot = (size_t)-1;
    std::vector<char> host;
    int rc = build_uniforms(c, kind, uniforms, ubytes, nviews, host);
    if (rc || host.empty()) return rc;
    cudaError_t e = cudaSuccess;
    void* d = c->arena.alloc(host.size(), e);
    CU(e);
    rc = param_copy(c, d, host.data(), host.size(), slot);   // `host` is pageable: staged (or kept by the recording) on return
    if (rc) return rc;
    *dev_out = d;
    return TRB_OK;
}
// [nviews][32] doubles: ModelView, Perspective of every view
void build_mats(const double* mv, const double* pr, int nv, std::vector<double>& hm) {
    hm.resize((size_t)32 * nv);
    for (int v = 0; v < nv; ++v) {
        memcpy(&hm[(size_t)v * 32], mv + 16 * v, 128);
        memcpy(&hm[(size_t)v * 32 + 16], pr + 16 * v, 128);
    }
}
// the LitF blocks of a lit mesh draw (fastshade.cuh); empty for the other shaders.  Handles were validated by build_uniforms.
void build_litf(TrbCtx* c, int kind, const void* uniforms, const double* mv, int nv, std::vector<trbf::LitF>& hl) {
    hl.clear();
    if (kind != TRB_SHADER_PHONG && kind != TRB_SHADER_EYE && kind != TRB_SHADER_SHADOW_PHONG) return;
    hl.resize(nv);
    for (int v = 0; v < nv; ++v) {
        const TrbPhongUniforms& u = kind == TRB_SHADER_SHADOW_PHONG ? ((const TrbShadowUniforms*)uniforms)[v].phong
                                                                    : ((const TrbPhongUniforms*)uniforms)[v];
        trbf::LitF& L = hl[v];
        for (int i = 0; i < 12; ++i) L.mv[i] = (float)mv[16 * v + i];
        L.key = trbf::F3{(float)u.key_dir_eye[0], (float)u.key_dir_eye[1], (float)u.key_dir_eye[2]};
        L.fill = trbf::F3{(float)u.fill_dir_eye[0], (float)u.fill_dir_eye[1], (float)u.fill_dir_eye[2]};
        L.rim = trbf::F3{(float)u.rim_dir_eye[0], (float)u.rim_dir_eye[1], (float)u.rim_dir_eye[2]};
        L.normal_map_strength = (float)u.normal_map_strength;
        // the maps the lit pixel samples; the specular map never changes a pixel (fastshade.cuh) and is not passed on
        L.dbpp = L.nbpp = L.dw = L.dh = L.nw = L.nh = 0;
        L.diffuse = L.normal = nullptr;
        if (u.diffuse) {
            const Tex& t = c->textures[u.diffuse - 1];
            L.diffuse = t.px; L.dw = (uint32_t)t.w; L.dh = (uint32_t)t.h; L.dbpp = (uint32_t)t.bpp;
        }
        if (u.normal) {
            const Tex& t = c->textures[u.normal - 1];
            L.normal = t.px; L.nw = (uint32_t)t.w; L.nh = (uint32_t)t.h; L.nbpp = (uint32_t)t.bpp;
        }
    }
}

int tga_finish(TrbCtx* c, TrbCtx::TgaJob& j, bool wait_copies);   // asynchronous TGA writer, defined with trb_encode_tga_async

constexpr unsigned long long COMM_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;   // a peer that stays silent for 20 s is gone

int comm_check_timeout(TrbCtx* c) {
    if (c->host_total[8]) return fail(c, TRB_E_COMM, "composite group: a peer did not reach the frame within 20 s");
    return TRB_OK;
}
// make this rank's stream wait until every peer has published `drawn` (done == false) or `done` (true) for frame comm.seq
int comm_wait_peers(TrbCtx* c, bool done) {
    TrbCtx::Comm& m = c->comm;
    if (m.seq == 0) return TRB_OK;
    int rc = comm_check_timeout(c);
    if (rc) return rc;
    if (m.in_process && m.ev_drawn) {            // events: safe for several contexts on ONE device
        for (int r = 0; r < m.n; ++r)
            if (r != m.rank) CU(cudaStreamWaitEvent(c->stream, done ? m.members[r]->comm.ev_done : m.members[r]->comm.ev_drawn, 0));
        return TRB_OK;
    }
    CommWait w;
    w.n = 0;
    for (int r = 0; r < m.n; ++r)
        if (r != m.rank) w.flag[w.n++] = done ? &m.peer_flags[r]->done : &m.peer_flags[r]->drawn;
    if (w.n == 0) return TRB_OK;
    Launch L(c, "k_comm_wait");
    k_comm_wait<<<1, 1, 0, c->stream>>>(w, m.seq, COMM_TIMEOUT_NS, c->host_total_dev + 8);
    CU(cudaGetLastError());
    return TRB_OK;
}
// the fused composite + shade of rows [y0, y1) queued on the context's stream; no synchronisation
int enqueue_composite_shade(TrbCtx* c, int y0, int y1) {
    if (snapshot_window(c)) {               // the composite rewrites keys of pixels no bin of this rank knows about
        int rs = snapshot_save_tiles(c, nullptr, nullptr);
        if (rs) return rs;
    }
    // the local planes may have been reallocated since the peers were opened
    c->peers.key[c->peer_rank] = c->frame.zkey;
    c->peers.vis[c->peer_rank] = c->frame.vis;
    if (y1 > y0 && !c->draws.empty()) {
        size_t bytes = c->draws.size() * sizeof(DrawDev);
        CU(c->draw_table.ensure(bytes, c->stream));
        CU(cudaMemcpyAsync(c->draw_table.p, c->draws.data(), bytes, cudaMemcpyHostToDevice, c->stream));
        bool config2 = false;
        for (const DrawDev& d : c->draws) config2 |= d.kind >= 4;
        // one CTA per chunk of the "touched" maps that overlaps the rows [y0, y1)
        const unsigned long long p_first = (unsigned long long)y0 * c->frame.W, p_last = (unsigned long long)y1 * c->frame.W;
        const unsigned long long n = ((p_last - 1) / COMPOSITE_CHUNK - p_first / COMPOSITE_CHUNK + 1) * COMPOSITE_CHUNK;
        Launch L(c, "k_composite_shade_p2p");
        const DrawDev* table = c->draw_table.as<DrawDev>();
        const int nd = (int)c->draws.size();
        bool fast = !c->shade_exact;
        for (const DrawDev& d : c->draws) fast &= !((d.kind == 1 || d.kind == 2 || d.kind == 4) && !d.litf);
        switch ((config2 ? 2 : 0) | (fast ? 1 : 0)) {
            case 0: k_composite_shade_p2p<false, false><<<blocks_for(n), TPB, 0, c->stream>>>(c->frame, c->peers, table, nd, y0, y1); break;
            case 1: k_composite_shade_p2p<false, true><<<blocks_for(n), TPB, 0, c->stream>>>(c->frame, c->peers, table, nd, y0, y1); break;
            case 2: k_composite_shade_p2p<true, false><<<blocks_for(n), TPB, 0, c->stream>>>(c->frame, c->peers, table, nd, y0, y1); break;
            default: k_composite_shade_p2p<true, true><<<blocks_for(n), TPB, 0, c->stream>>>(c->frame, c->peers, table, nd, y0, y1); break;
        }
    }
    CU(cudaGetLastError());
    c->draws.clear();      // the pageable table was staged before cudaMemcpyAsync returned
    return TRB_OK;
}
void comm_rows(int H, int rank, int n, int* y0, int* y1) {
    const int base = H / n, rem = H % n;
    *y0 = rank * base + std::min(rank, rem);
    *y1 = *y0 + base + (rank < rem ? 1 : 0);
}

void ssao_dirs(SsaoDirs& d) {
    for (int k = 0; k < 8; ++k) {
        double angle = 2.0 * M_PI * k / 8;  // main.cpp:333, host libm like the reference
        d.dx[k] = cos(angle);
        d.dy[k] = sin(angle);
    }
}

}  // namespace

extern "C" {

int trb_create(int device, TrbCtx** out) {
    if (!out) return TRB_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return TRB_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TRB_E_CUDA;
    if (prop.major != 10) return TRB_E_CUDA;  // the library only carries sm_100a code
    if (cudaSetDevice(device) != cudaSuccess) return TRB_E_CUDA;
    TrbCtx* c = new TrbCtx();
    c->device = device;
    c->sms = prop.multiProcessorCount;   // 148 on B200: persistent grids and grid caps are sized from it
    if (const char* e = getenv("TRB_BIG_NS")) c->big_ns = std::max(1, atoi(e));
    if (const char* e = getenv("TRB_SMALL_MIN")) c->small_min = std::max(1, atoi(e));
    if (const char* e = getenv("TRB_LARGE_NS")) c->large_ns = std::max(1, atoi(e));
    if (const char* e = getenv("TRB_DIRECT_AREA")) c->direct_area = std::max(0, atoi(e));
    if (const char* e = getenv("TRB_DIRECT_BY_PIXEL")) c->direct_by_pixel = atoi(e) != 0;
    if (const char* e = getenv("TRB_WARP_MAX")) c->warp_max = (uint32_t)std::max(0, atoi(e));
    if (const char* e = getenv("TRB_RW_BLOCKS")) c->rw_blocks = atoi(e);
    if (const char* e = getenv("TRB_SHADE_EXACT")) c->shade_exact = atoi(e) != 0;
    if (const char* e = getenv("TRB_TMA")) c->use_tma = atoi(e) != 0;
    if (const char* e = getenv("TRB_LAZY_SNAPSHOT")) c->lazy_snapshot = atoi(e) != 0;
    if (const char* e = getenv("TRB_COLLECT_BY_TILES")) c->collect_by_tiles = atoi(e) != 0;
    if (const char* e = getenv("TRB_LAZY_SNAPSHOT_MAX_TRIS")) c->lazy_max_tris = (uint64_t)std::max(0ll, atoll(e));
    if (const char* e = getenv("TRB_SYNC_DRAWS")) c->sync_draws = atoi(e) != 0;
    if (const char* e = getenv("TRB_BIN_CAP")) c->bin_cap_fixed = (uint32_t)std::max(1, atoi(e));
    if (const char* e = getenv("TRB_SPLIT")) c->split_on = atoi(e) != 0;
    if (const char* e = getenv("TRB_SPLIT_S")) c->split_s = (uint32_t)std::min(65535, std::max(32, atoi(e)));
    if (const char* e = getenv("TRB_SPLIT_CAP")) c->split_cap_fixed = (uint32_t)std::max(1, atoi(e));
    if (const char* e = getenv("TRB_SHARD_SHIFT")) c->shard_shift = (uint32_t)std::min(24, std::max(5, atoi(e)));
    if (const char* e = getenv("TRB_MESH_ORDER_MIN_TRIS")) c->order_min_tris = (uint64_t)std::max(0ll, atoll(e));
    if (const char* e = getenv("TRB_MESH_ORDER_VERTICES")) c->order_vertices = atoi(e) != 0;
    if (const char* e = getenv("TRB_SHARE_VERTEX")) c->share_vertex = atoi(e) != 0;
    if (const char* e = getenv("TRB_SHARE_VERTEX_MIN_RANKS")) c->share_vertex_min_ranks = std::max(2, atoi(e));
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->upload_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&c->ev_a) != cudaSuccess || cudaEventCreate(&c->ev_b) != cudaSuccess ||
        cudaHostAlloc((void**)&c->host_total, 64, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&c->host_total_dev, c->host_total, 0) != cudaSuccess) {
        delete c;
        return TRB_E_CUDA;
    }
    memset(c->host_total, 0, 64);
    // 9 CTAs x 24 KB of per-warp tiles per SM: ask for the large shared-memory carveout
    cudaFuncSetAttribute(k_raster_warp<6, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_raster_warp<7, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_raster_warp<8, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_raster_warp<6, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    *out = c;
    return TRB_OK;
}

int trb_destroy(TrbCtx* c) {
    if (!c) return TRB_E_ARG;
    cudaSetDevice(c->device);
    if (c->rec) {                         // destroyed in the middle of a recording: end the capture first
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(c->stream, &g);
        c->rec->graph = g;
        recording_destroy(c->rec);
        c->rec = nullptr;
        if (t_recording > 0) --t_recording;
        (void)cudaGetLastError();
    }
    cudaStreamSynchronize(c->stream);
    if (c->upload_stream) cudaStreamSynchronize(c->upload_stream);
    for (auto& m : c->meshes)
        if (m.alive) {
            cudaFree(m.pos4);
            cudaFree(m.attr8);
            if (m.idx) cudaFree(m.idx);
            if (m.perm) cudaFree(m.perm);
            if (m.idx_perm) cudaFree(m.idx_perm);
            if (m.inv_perm) cudaFree(m.inv_perm);
            if (m.vmark) cudaFree(m.vmark);
        }
    for (auto& t : c->textures)
        if (t.alive) cudaFree(t.px);
    DevBuf* bufs[] = {&c->zkey, &c->vis, &c->color, &c->stats, &c->zsnap, &c->zlocal, &c->snap_saved, &c->draw_table, &c->shade_list, &c->tribox, &c->trirec,
                      &c->drawzero, &c->offsets, &c->bins, &c->scan_sums, &c->scan_total, &c->heavy_list, &c->split_items, &c->split_keys, &c->split_ids, &c->direct_list, &c->rle_work, &c->rle_src, &c->rle_out, &c->scratch_a,
                      &c->scratch_b};
    for (DevBuf* b : bufs) b->release();
    for (auto& b : c->shadow_maps) b.keys.release();
    for (auto& b : c->shadow_pool) b.release();
    trb_comm_close(c);
    trb_ipc_close_peers(c);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        for (int i = 0; i < TrbCtx::RB_SLOTS; ++i) {
            c->rb[i].release();
            cudaEventDestroy(c->rb_ready[i]);
            cudaEventDestroy(c->rb_done[i]);
        }
        cudaStreamDestroy(c->copy_stream);
    }
    for (auto& j : c->tga) {
        j.out.release();
        if (j.offs_host) cudaFreeHost(j.offs_host);
        if (j.offs_ready) cudaEventDestroy(j.offs_ready);
        if (j.done) cudaEventDestroy(j.done);
    }
    for (Recording* r : c->recordings) recording_destroy(r);
    c->arena.release();
    c->cache.release();
    c->ring.release();
    if (c->upload_ev) cudaEventDestroy(c->upload_ev);
    if (c->upload_stream) cudaStreamDestroy(c->upload_stream);
    for (auto& p : c->prof_pending) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    cudaEventDestroy(c->ev_a);
    cudaEventDestroy(c->ev_b);
    cudaFreeHost(c->host_total);
    cudaStreamDestroy(c->stream);
    delete c;
    return TRB_OK;
}

const char* trb_last_error(TrbCtx* c) { return c ? c->err.c_str() : "null context"; }
const char* trb_backend_name(void) { return "cuda-sm100a"; }

}  // extern "C"

namespace {
// device block for a resource + `bytes` of it filled through the pinned ring on the upload stream.
// fill(dst, offset, n) writes bytes [offset, offset + n) of the device layout into pinned memory.
template <class Fill>
int upload_block(TrbCtx* c, void** dev, size_t bytes, size_t granule, Fill fill) {
    cudaEvent_t wait_for = nullptr;
    CU(c->cache.get(dev, bytes, &wait_for, c->upload_stream));
    if (wait_for) CU(cudaStreamWaitEvent(c->upload_stream, wait_for, 0));   // its last readers (render stream) first
    const size_t step = UploadRing::CHUNK / granule * granule;
    for (size_t off = 0; off < bytes; off += step) {
        const size_t n = std::min(step, bytes - off);
        char* stage = nullptr;
        cudaEvent_t done = nullptr;
        CU(c->ring.slot(&stage, &done));
        fill(stage, off, n);
        CU(cudaMemcpyAsync((char*)*dev + off, stage, n, cudaMemcpyHostToDevice, c->upload_stream));
        CU(cudaEventRecord(done, c->upload_stream));
    }
    return TRB_OK;
}
// largest vertex index of a mesh (upload_mesh validates the range before anything reaches the device);
// written as a reduction so that the AVX2 clone runs at memory speed on multi-million-triangle meshes
__attribute__((target("avx2"))) uint32_t max_index_avx2(const uint32_t* idx, uint64_t n) {
    uint32_t m = 0;
    for (uint64_t i = 0; i < n; ++i) m = idx[i] > m ? idx[i] : m;
    return m;
}
uint32_t max_index(const uint32_t* idx, uint64_t n) {
    if (__builtin_cpu_supports("avx2")) return max_index_avx2(idx, n);
    uint32_t m = 0;
    for (uint64_t i = 0; i < n; ++i) m = idx[i] > m ? idx[i] : m;
    return m;
}
// page-locked host memory can be handed to the copy engine as it is (cudaHostAlloc / cudaHostRegister /
// torch pin_memory): no staging copy on the CPU.  The caller keeps such arrays alive and unchanged
// until the next synchronising call (trb.h).
bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
// device block filled by one DMA straight from pinned host memory
int upload_pinned(TrbCtx* c, void** dev, const void* src, size_t bytes) {
    cudaEvent_t wait_for = nullptr;
    CU(c->cache.get(dev, bytes, &wait_for, c->upload_stream));
    if (wait_for) CU(cudaStreamWaitEvent(c->upload_stream, wait_for, 0));
    CU(cudaMemcpyAsync(*dev, src, bytes, cudaMemcpyHostToDevice, c->upload_stream));
    return TRB_OK;
}
// processing order of a freshly uploaded large indexed mesh, queued on the upload stream behind its arrays
int build_mesh_order(TrbCtx* c, Mesh& m) {
    const uint64_t ntris = m.nidx / 3;
    if (c->order_min_tris == 0 || ntris < c->order_min_tris || ntris > 0x55555555ull || ntris == 0) return TRB_OK;
    if (m.idx && c->order_vertices && m.nverts <= 0x7fffffffu) {
        // indexed mesh: first the vertices into Morton numbering (second copies of the arrays, the first ones are recycled;
        // the index buffer is rewritten in place)
        const size_t vbytes = trb_vertex_order_scratch_bytes(m.nverts);
        void* vs = nullptr;
        float4* pos2 = nullptr;
        float* attr2 = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
        CU(c->cache.get(&vs, vbytes, &e0, c->upload_stream));
        CU(c->cache.get((void**)&pos2, (size_t)m.nverts * 16, &e1, c->upload_stream));
        CU(c->cache.get((void**)&attr2, (size_t)m.nverts * 32, &e2, c->upload_stream));
        if (e0) CU(cudaStreamWaitEvent(c->upload_stream, e0, 0));
        if (e1) CU(cudaStreamWaitEvent(c->upload_stream, e1, 0));
        if (e2) CU(cudaStreamWaitEvent(c->upload_stream, e2, 0));
        {
            Launch L(c, "vertex_order", c->upload_stream, /*kernel=*/false);
            CU(trb_vertex_order_apply(m.pos4, m.attr8, m.nverts, m.idx, m.nidx, pos2, attr2, vs, vbytes, c->sms, c->upload_stream));
        }
        c->cache.put(vs, vbytes, c->upload_stream);
        c->cache.put(m.pos4, (size_t)m.nverts * 16, c->upload_stream);
        c->cache.put(m.attr8, (size_t)m.nverts * 32, c->upload_stream);
        m.pos4 = pos2;
        m.attr8 = attr2;
    }
    const size_t sbytes = trb_mesh_order_scratch_bytes((uint32_t)ntris);
    void* scratch = nullptr;
    cudaEvent_t w0 = nullptr, w1 = nullptr, w2 = nullptr;
    CU(c->cache.get((void**)&m.perm, ntris * 4, &w0, c->upload_stream));
    if (m.idx) CU(c->cache.get((void**)&m.idx_perm, m.nidx * 4, &w1, c->upload_stream));
    CU(c->cache.get(&scratch, sbytes, &w2, c->upload_stream));
    if (w0) CU(cudaStreamWaitEvent(c->upload_stream, w0, 0));
    if (w1) CU(cudaStreamWaitEvent(c->upload_stream, w1, 0));
    if (w2) CU(cudaStreamWaitEvent(c->upload_stream, w2, 0));
    {
        Launch L(c, "mesh_order", c->upload_stream, /*kernel=*/false);   // several kernels incl. the library sort: timed as one span
        CU(trb_mesh_order_build(m.pos4, m.nverts, m.idx, (uint32_t)ntris, m.perm, m.idx_perm, scratch, sbytes, c->sms, c->upload_stream));
    }
    c->cache.put(scratch, sbytes, c->upload_stream);
    if (!m.idx) {
        // soup: the vertex arrays go into processing order themselves (second copies, the first ones are recycled)
        float4* pos2 = nullptr;
        float* attr2 = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
        CU(c->cache.get((void**)&pos2, (size_t)m.nverts * 16, &e0, c->upload_stream));
        CU(c->cache.get((void**)&attr2, (size_t)m.nverts * 32, &e1, c->upload_stream));
        CU(c->cache.get((void**)&m.inv_perm, ntris * 4, &e2, c->upload_stream));
        if (e0) CU(cudaStreamWaitEvent(c->upload_stream, e0, 0));
        if (e1) CU(cudaStreamWaitEvent(c->upload_stream, e1, 0));
        if (e2) CU(cudaStreamWaitEvent(c->upload_stream, e2, 0));
        {
            Launch L(c, "soup_order", c->upload_stream, /*kernel=*/false);
            CU(trb_soup_order_apply(m.pos4, m.attr8, m.nverts, m.perm, (uint32_t)ntris, pos2, attr2, m.inv_perm, c->upload_stream));
        }
        c->cache.put(m.pos4, (size_t)m.nverts * 16, c->upload_stream);
        c->cache.put(m.attr8, (size_t)m.nverts * 32, c->upload_stream);
        m.pos4 = pos2;
        m.attr8 = attr2;
    }
    return TRB_OK;
}
// everything uploaded so far becomes visible to the render stream
int publish_uploads(TrbCtx* c) {
    CU(cudaEventRecord(c->upload_ev, c->upload_stream));
    CU(cudaStreamWaitEvent(c->stream, c->upload_ev, 0));
    return TRB_OK;
}
}  // namespace

extern "C" {

int trb_upload_mesh(TrbCtx* c, const float* pos3, const float* nrm3, const float* uv2, uint32_t nverts,
                    const uint32_t* idx, uint64_t nidx, TrbMesh* out) {
    NOT_WHILE_RECORDING(c, "upload_mesh");
    HostSpan host_span_("trb_upload_mesh");
    if (!c || !pos3 || !out || nidx % 3 || nverts == 0) return fail(c, TRB_E_ARG, "upload_mesh: bad argument");
    if (nidx / 3 >= 0xFFFFFFF0ull) return fail(c, TRB_E_ARG, "upload_mesh: too many triangles");
    int rc = check_device(c);
    if (rc) return rc;
    if (idx && nidx && max_index(idx, nidx) >= nverts) return fail(c, TRB_E_ARG, "upload_mesh: index out of range");
    if (!idx && nidx > nverts) return fail(c, TRB_E_ARG, "upload_mesh: implicit indices exceed nverts");
    // The host interleaves into the two device layouts - float4 positions for the coalesced vertex
    // kernel, 32-byte {pos,nrm,uv} records (one DRAM sector) for the shade kernel's gathers - straight
    // into pinned chunks that the upload stream drains while the render stream keeps rendering.
    // The caller's arrays are fully consumed when this returns.
    Mesh m;
    m.nverts = nverts;
    m.nidx = nidx;
    if (is_pinned(pos3) && (!nrm3 || is_pinned(nrm3)) && (!uv2 || is_pinned(uv2)) && (!idx || is_pinned(idx))) {
        // pinned caller: raw arrays by DMA, interleaved by a kernel on the upload stream - no CPU copy at all
        float *rp = nullptr, *rn = nullptr, *ru = nullptr;
        rc = upload_pinned(c, (void**)&rp, pos3, (size_t)nverts * 12);
        if (!rc && nrm3) rc = upload_pinned(c, (void**)&rn, nrm3, (size_t)nverts * 12);
        if (!rc && uv2) rc = upload_pinned(c, (void**)&ru, uv2, (size_t)nverts * 8);
        if (!rc && idx) rc = upload_pinned(c, (void**)&m.idx, idx, nidx * 4);
        if (rc) return rc;
        cudaEvent_t w0 = nullptr, w1 = nullptr;
        CU(c->cache.get((void**)&m.pos4, (size_t)nverts * 16, &w0, c->upload_stream));
        CU(c->cache.get((void**)&m.attr8, (size_t)nverts * 32, &w1, c->upload_stream));
        if (w0) CU(cudaStreamWaitEvent(c->upload_stream, w0, 0));
        if (w1) CU(cudaStreamWaitEvent(c->upload_stream, w1, 0));
        {
            Launch L(c, "k_interleave_mesh", c->upload_stream);
            k_interleave_mesh<<<blocks_for(nverts), TPB, 0, c->upload_stream>>>(rp, rn, ru, nverts, m.pos4, m.attr8);
        }
        CU(cudaGetLastError());
        c->cache.put(rp, (size_t)nverts * 12, c->upload_stream);   // raw arrays: free again once the kernel has run
        c->cache.put(rn, (size_t)nverts * 12, c->upload_stream);
        c->cache.put(ru, (size_t)nverts * 8, c->upload_stream);
        rc = build_mesh_order(c, m);
        if (rc) return rc;
        rc = publish_uploads(c);
        if (rc) return rc;
        m.alive = true;
        size_t slot = 0;
        while (slot < c->meshes.size() && c->meshes[slot].alive) ++slot;
        if (slot == c->meshes.size()) c->meshes.push_back(m); else c->meshes[slot] = m;
        *out = slot + 1;
        return TRB_OK;
    }
    rc = upload_block(c, (void**)&m.pos4, (size_t)nverts * 16, 16, [&](char* dst, size_t off, size_t n) {
        float* p = reinterpret_cast<float*>(dst);
        const size_t v0 = off / 16, nv = n / 16;
        for (size_t v = 0; v < nv; ++v, p += 4) {
            const float* s = pos3 + 3 * (v0 + v);
            p[0] = s[0]; p[1] = s[1]; p[2] = s[2]; p[3] = 1.0f;
        }
    });
    if (rc) return rc;
    rc = upload_block(c, (void**)&m.attr8, (size_t)nverts * 32, 32, [&](char* dst, size_t off, size_t n) {
        float* a = reinterpret_cast<float*>(dst);
        const size_t v0 = off / 32, nv = n / 32;
        for (size_t v = 0; v < nv; ++v, a += 8) {
            const size_t g = v0 + v;
            a[0] = pos3[3 * g]; a[1] = pos3[3 * g + 1]; a[2] = pos3[3 * g + 2];
            a[3] = nrm3 ? nrm3[3 * g] : 0.f;           // Model::normal fallback (0,0,1), model.cpp:404
            a[4] = nrm3 ? nrm3[3 * g + 1] : 0.f;
            a[5] = nrm3 ? nrm3[3 * g + 2] : 1.f;
            a[6] = uv2 ? uv2[2 * g] : 0.f;
            a[7] = uv2 ? uv2[2 * g + 1] : 0.f;
        }
    });
    if (rc) return rc;
    if (idx) {
        rc = upload_block(c, (void**)&m.idx, nidx * 4, 4, [&](char* dst, size_t off, size_t n) {
            memcpy(dst, reinterpret_cast<const char*>(idx) + off, n);
        });
        if (rc) return rc;
    }
    rc = build_mesh_order(c, m);
    if (rc) return rc;
    rc = publish_uploads(c);
    if (rc) return rc;
    m.alive = true;
    size_t slot = 0;
    while (slot < c->meshes.size() && c->meshes[slot].alive) ++slot;  // handles of freed meshes are reused
    if (slot == c->meshes.size()) c->meshes.push_back(m); else c->meshes[slot] = m;
    *out = slot + 1;
    return TRB_OK;
}

int trb_free_mesh(TrbCtx* c, TrbMesh h) {
    NOT_WHILE_RECORDING(c, "free_mesh");
    HostSpan host_span_("trb_free_mesh");
    if (!c || h == 0 || h > c->meshes.size() || !c->meshes[h - 1].alive) return fail(c, TRB_E_ARG, "free_mesh");
    int rc = check_device(c);
    if (rc) return rc;
    // an unflushed draw still points at this mesh (the deferred shade kernels read idx / attr8 at the next
    // flush): shade first, so the event the block is tagged with covers its last reader
    if (!c->draws.empty() && (rc = do_flush(c))) return rc;
    Mesh& m = c->meshes[h - 1];
    ++g_generation;                    // recordings that draw it are stale now
    c->cache.put(m.pos4, (size_t)m.nverts * 16, c->stream);   // tagged: reusable once the kernels queued so far are done
    c->cache.put(m.attr8, (size_t)m.nverts * 32, c->stream);
    if (m.idx) c->cache.put(m.idx, m.nidx * 4, c->stream);
    if (m.perm) c->cache.put(m.perm, m.nidx / 3 * 4, c->stream);
    if (m.idx_perm) c->cache.put(m.idx_perm, m.nidx * 4, c->stream);
    if (m.inv_perm) c->cache.put(m.inv_perm, m.nidx / 3 * 4, c->stream);
    if (m.vmark) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(m.vmark);
    }
    m = Mesh();
    return TRB_OK;
}

int trb_upload_texture(TrbCtx* c, const uint8_t* texels, int w, int h, int bpp, TrbTex* out) {
    NOT_WHILE_RECORDING(c, "upload_texture");
    HostSpan host_span_("trb_upload_texture");
    // 65535: the TGA header's 16-bit sizes (tgaimage.h); it also keeps every texel index below 2^32
    if (!c || !texels || !out || w <= 0 || h <= 0 || w > 65535 || h > 65535 || (bpp != 1 && bpp != 3 && bpp != 4))
        return fail(c, TRB_E_ARG, "upload_texture: bad argument");
    int rc = check_device(c);
    if (rc) return rc;
    Tex t;
    t.w = w;
    t.h = h;
    t.bpp = bpp;
    size_t bytes = (size_t)w * h * bpp;
    if (is_pinned(texels))
        rc = upload_pinned(c, (void**)&t.px, texels, bytes);
    else
        rc = upload_block(c, (void**)&t.px, bytes, 1, [&](char* dst, size_t off, size_t n) { memcpy(dst, texels + off, n); });
    if (rc) return rc;
    rc = publish_uploads(c);
    if (rc) return rc;
    t.alive = true;
    size_t slot = 0;
    while (slot < c->textures.size() && c->textures[slot].alive) ++slot;
    if (slot == c->textures.size()) c->textures.push_back(t); else c->textures[slot] = t;
    *out = slot + 1;
    return TRB_OK;
}

int trb_free_texture(TrbCtx* c, TrbTex h) {
    NOT_WHILE_RECORDING(c, "free_texture");
    HostSpan host_span_("trb_free_texture");
    if (!c || h == 0 || h > c->textures.size() || !c->textures[h - 1].alive) return fail(c, TRB_E_ARG, "free_texture");
    int rc = check_device(c);
    if (rc) return rc;
    if (!c->draws.empty() && (rc = do_flush(c))) return rc;   // pending draws sample it at the next flush (see free_mesh)
    ++g_generation;                    // recordings that sample it are stale now
    Tex& x = c->textures[h - 1];
    c->cache.put(x.px, (size_t)x.w * x.h * x.bpp, c->stream);
    x = Tex();
    return TRB_OK;
}

int trb_begin_batch(TrbCtx* c, int w, int h, int nviews) {
    HostSpan host_span_("trb_begin_batch");
    if (!c || w <= 0 || h <= 0 || nviews <= 0 || nviews > 65535 || w > 65536 || h > 65536)
        return fail(c, TRB_E_ARG, "begin_batch: bad size");
    int rc = check_device(c);
    if (rc) return rc;
    FrameDev& f = c->frame;
    f.W = w;
    f.H = h;
    f.tw = (w + TILE - 1) / TILE;
    f.th = (h + TILE - 1) / TILE;
    f.ntiles = f.tw * f.th;
    f.nviews = nviews;
    f.npix = (unsigned long long)w * h;
    const size_t total = (size_t)f.npix * nviews;
    CU(c->zkey.ensure(total * 8, c->stream));
    CU(c->vis.ensure(total * 4, c->stream));
    CU(c->color.ensure(total * 3, c->stream));
    CU(c->stats.ensure(sizeof(DevStats) * nviews, c->stream));
    f.zkey = c->zkey.as<unsigned long long>();
    f.vis = c->vis.as<uint32_t>();
    f.color = c->color.as<uint8_t>();
    f.stats = c->stats.as<DevStats>();
    for (int i = 0; i < 16; ++i) f.viewport[i] = (i % 5 == 0) ? 1.0 : 0.0;
    c->draws.clear();
    c->arena.reset();
    ++c->cache.frame;
    c->next_id = 0;
    c->foreign_ids = false;
    c->tris_submitted = 0;
    c->have_snapshot = false;
    c->snap_stale = false;
    c->snap_lazy = c->snap_all = false;
    c->shade_row0 = 0;
    c->shade_row1 = -1;
    if (c->comm.n > 0) {
        // peers read this rank's planes during their composite: the planes must be the ones they opened, and the clear
        // below must not start before every peer has finished reading the previous frame
        if (nviews != 1 || w != c->comm.W || h != c->comm.H || c->zkey.p != c->comm.key_at_init)
            return fail(c, TRB_E_COMM, "begin_frame: the frame differs from the one the composite group was opened with "
                                       "(trb_comm_close, then open the group again)");
        rc = comm_wait_peers(c, /*done=*/true);
        if (rc) return rc;
    }
    {
        unsigned grid = (unsigned)std::min<unsigned long long>(blocks_for(total), (unsigned long long)c->sms * 32);
        Launch L(c, "k_clear");
        k_clear<<<grid, TPB, 0, c->stream>>>(f, c->clear[0], c->clear[1], c->clear[2]);
    }
    CU(cudaGetLastError());
    c->in_frame = true;
    return TRB_OK;
}
int trb_begin_frame(TrbCtx* c, int w, int h) { return trb_begin_batch(c, w, h, 1); }

int trb_set_clear_color(TrbCtx* c, uint8_t b, uint8_t g, uint8_t r) {
    if (!c) return TRB_E_ARG;
    c->clear[0] = b;
    c->clear[1] = g;
    c->clear[2] = r;
    return TRB_OK;
}
int trb_set_viewport(TrbCtx* c, const double* v) {
    if (!c || !v) return fail(c, TRB_E_ARG, "set_viewport");
    memcpy(c->frame.viewport, v, sizeof(double) * 16);
    return TRB_OK;
}

}  // extern "C"

namespace {

// One mesh draw.  shard_n <= 1: the triangle range [first_tri, first_tri + ntris) of the index buffer.
// shard_n > 1 (trb_draw_shard): rank shard_r's share of the WHOLE mesh - first_tri / ntris are ignored.
int draw_mesh(TrbCtx* c, TrbMesh mesh, const double* mv, const double* pr, int kind, const void* uniforms,
              size_t ubytes, uint64_t first_tri, uint64_t ntris, uint32_t shard_n, uint32_t shard_r) {
    if (!c || !c->in_frame) return fail(c, TRB_E_ARG, "draw: no frame");
    if (!mv || !pr) return fail(c, TRB_E_ARG, "draw: null matrix");
    if (mesh == 0 || mesh > c->meshes.size() || !c->meshes[mesh - 1].alive) return fail(c, TRB_E_ARG, "draw: bad mesh");
    const Mesh& m = c->meshes[mesh - 1];
    const uint64_t mesh_tris = m.nidx / 3;
    // ids of a sharded draw are those of the whole mesh on every rank: base + triangle + 1
    const uint64_t id_base0 = c->next_id;
    uint64_t id_span = ntris;               // ids this draw consumes
    uint32_t share_slots = 0;               // shard in processing-order space: slots of this rank
    bool shard_order = false;
    if (shard_n > 1) {
        if (shard_r >= shard_n) return fail(c, TRB_E_ARG, "draw_shard: rank outside the shard count");
        id_span = mesh_tris;
        c->foreign_ids = true;
        if (m.perm) {
            // blocks b = shard_r (mod shard_n) of 2^SHARD_SHIFT positions of the mesh's processing order
            const uint64_t B = 1ull << c->shard_shift, nblocks = (mesh_tris + B - 1) / B;
            const uint64_t mine = nblocks > shard_r ? (nblocks - shard_r + shard_n - 1) / shard_n : 0;
            uint64_t slots = mine * B;
            if (mine && (nblocks - 1) % shard_n == shard_r) slots -= nblocks * B - mesh_tris;   // the last block is partial
            share_slots = (uint32_t)slots;
            shard_order = true;
            first_tri = 0;
            ntris = slots;
        } else {
            // small mesh (no processing order): a contiguous range of the index buffer
            const uint64_t base = mesh_tris / shard_n, rem = mesh_tris % shard_n;
            first_tri = shard_r * base + std::min<uint64_t>(shard_r, rem);
            ntris = base + (shard_r < rem ? 1 : 0);
            c->next_id = id_base0 + first_tri;
        }
    }
    {   // no wrap-around: both values end up below the mesh's 0xFFFFFFF0 triangle limit
        const uint64_t mt = mesh_tris;
        if (first_tri > mt || ntris > mt - first_tri) return fail(c, TRB_E_ARG, "draw: triangle range");
    }
    if (id_base0 + id_span >= 0xFFFFFFF0ull) return fail(c, TRB_E_ARG, "draw: triangle id space exhausted");
    int rc = check_device(c);
    if (rc) return rc;
    const int nv = c->frame.nviews;
    c->tris_submitted += ntris;
    if (ntris == 0) {
        c->next_id = id_base0 + id_span;
        if (c->rec) c->rec->draws.push_back(Recording::Draw{kind, nv, (size_t)-1, (size_t)-1, (size_t)-1, ubytes, {}, {}});
        return TRB_OK;
    }
    const void* dun = nullptr;
    Recording::Draw slots{kind, nv, (size_t)-1, (size_t)-1, (size_t)-1, ubytes, {}, {}};
    rc = resolve_uniforms(c, kind, uniforms, ubytes, nv, &dun, &slots.uni_off);
    if (rc) return rc;
    cudaError_t e = cudaSuccess;
    double* mats = (double*)c->arena.alloc(sizeof(double) * 32 * nv, e);
    CU(e);
    std::vector<double> hm;
    build_mats(mv, pr, nv, hm);
    rc = param_copy(c, mats, hm.data(), hm.size() * 8, &slots.mats_off);
    if (rc) return rc;
    const void* litf = nullptr;
    std::vector<trbf::LitF> hl;
    build_litf(c, kind, uniforms, mv, nv, hl);
    if (!hl.empty()) {
        void* dl = c->arena.alloc(sizeof(trbf::LitF) * nv, e);
        CU(e);
        rc = param_copy(c, dl, hl.data(), sizeof(trbf::LitF) * nv, &slots.litf_off);
        if (rc) return rc;
        litf = dl;
    }
    if (c->rec) {      // what the caller passed, kept so that a replay can change the matrices or the uniforms alone
        slots.mv.assign(mv, mv + (size_t)16 * nv);
        if (uniforms && ubytes) slots.uni.assign((const char*)uniforms, (const char*)uniforms + ubytes * nv);
        c->rec->draws.push_back(slots);
    }
    VRec* vrec = (VRec*)c->arena.alloc(sizeof(VRec) * (size_t)m.nverts * nv, e);
    CU(e);
    // One rank's share of an ordered mesh inside a composite group: only the vertices the share refers to go through
    // the vertex stage (the composite's shade pass transforms the vertices of other ranks' winners itself, a few per
    // pixel row it owns, instead of every rank transforming every vertex).  Not for the NCCL composite, whose shade
    // pass is the ordinary one.
    const uint8_t* vmark = nullptr;
    if (shard_order && c->comm.n > 0 && m.idx_perm && c->share_vertex && shard_n >= (uint32_t)c->share_vertex_min_ranks) {
        Mesh& mm = c->meshes[mesh - 1];
        if (!mm.vmark || mm.vmark_n != shard_n || mm.vmark_r != shard_r || mm.vmark_shift != c->shard_shift) {
            if (!mm.vmark) {
                if (t_recording) return fail(c, TRB_E_ARG, "record: draw_shard: render the frame once before recording it");
                CU(cudaMalloc((void**)&mm.vmark, mm.nverts));
            }
            CU(cudaMemsetAsync(mm.vmark, 0, mm.nverts, c->stream));
            Launch L(c, "k_mark_share_vertices");
            k_mark_share_vertices<<<blocks_for(share_slots), TPB, 0, c->stream>>>(mm.idx_perm, (uint32_t)mesh_tris, share_slots, shard_n,
                                                                                 shard_r, c->shard_shift, mm.vmark);
            mm.vmark_n = shard_n; mm.vmark_r = shard_r; mm.vmark_shift = c->shard_shift;
        }
        vmark = mm.vmark;
    }
    {
        Launch L(c, "k_vertex_mesh");
        k_vertex_mesh<<<dim3(blocks_for(m.nverts), nv), TPB, 0, c->stream>>>(c->frame, m.pos4, m.nverts, mats, vrec, vmark);
    }
    CU(cudaGetLastError());
    GeomArgs g;
    g.idx = m.idx;
    g.first_tri = (uint32_t)first_tri;
    g.ntris = (uint32_t)ntris;
    g.nverts = m.nverts;
    g.id_base = (uint32_t)c->next_id;
    g.vrec = vrec;
    // the mesh's processing order, unless the range is a small part of the mesh (the draw visits every slot of the mesh)
    const bool ordered = shard_order || (m.perm && (m.inv_perm || ntris * 16 >= mesh_tris));   // an ordered soup has no other order left
    g.perm = ordered ? m.perm : nullptr;
    g.idx_perm = ordered ? m.idx_perm : nullptr;
    g.nslots = shard_order ? share_slots : ordered ? (uint32_t)mesh_tris : g.ntris;
    g.shard_n = shard_order ? shard_n : 0u;
    g.shard_r = shard_order ? shard_r : 0u;
    g.shard_shift = c->shard_shift;
    g.nperm = (uint32_t)mesh_tris;
    if (shard_order) g.ntris = (uint32_t)mesh_tris;   // every triangle of the order is "inside the range"; the slots pick the share
    rc = raster_draw(c, g);   // `hm`, `hl` are pageable: their copies were staged before cudaMemcpyAsync returned
    if (rc) return rc;
    DrawDev d{};
    d.id_base = g.id_base;
    d.ntris = g.ntris;
    d.first_tri = g.first_tri;
    d.nverts = g.nverts;
    d.idx = m.idx;
    d.inv_perm = m.inv_perm;
    d.vmark = vmark;
    d.pos4 = m.pos4;
    d.attr8 = m.attr8;
    d.vrec = vrec;
    d.mats = mats;
    d.uniforms = dun;
    d.varyings = nullptr;
    d.litf = litf;
    d.kind = kind;
    d.mesh_ntris = (uint32_t)(m.nidx / 3);
    d.mesh_id_base = (long long)g.id_base - (long long)g.first_tri;
    c->draws.push_back(d);
    c->next_id = id_base0 + id_span;
    return TRB_OK;
}
}  // namespace

extern "C" {

int trb_draw_batch(TrbCtx* c, TrbMesh mesh, const double* mv, const double* pr, int kind, const void* uniforms,
                   size_t ubytes, uint64_t first_tri, uint64_t ntris) {
    HostSpan host_span_("trb_draw_batch");
    return draw_mesh(c, mesh, mv, pr, kind, uniforms, ubytes, first_tri, ntris, 0, 0);
}

int trb_draw_shard(TrbCtx* c, TrbMesh mesh, const double* mv, const double* pr, int kind, const void* uniforms,
                   size_t ubytes, int shard_rank, int shard_count) {
    NOT_WHILE_RECORDING(c, "draw_shard");
    HostSpan host_span_("trb_draw_shard");
    if (c && c->in_frame && c->frame.nviews != 1) return fail(c, TRB_E_ARG, "draw_shard: needs a single-view frame");
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) return fail(c, TRB_E_ARG, "draw_shard: bad rank / count");
    if (c && mesh != 0 && mesh <= c->meshes.size() && c->meshes[mesh - 1].alive && shard_count == 1)
        return draw_mesh(c, mesh, mv, pr, kind, uniforms, ubytes, 0, c->meshes[mesh - 1].nidx / 3, 0, 0);
    return draw_mesh(c, mesh, mv, pr, kind, uniforms, ubytes, 0, 0, (uint32_t)shard_count, (uint32_t)shard_rank);
}

int trb_draw(TrbCtx* c, TrbMesh mesh, const double* mv, const double* pr, int kind, const void* uniforms,
             size_t ubytes, uint64_t first_tri, uint64_t ntris) {
    if (c && c->in_frame && c->frame.nviews != 1) return fail(c, TRB_E_ARG, "draw: batch frame needs draw_batch");
    return trb_draw_batch(c, mesh, mv, pr, kind, uniforms, ubytes, first_tri, ntris);
}

int trb_submit_clip_triangles(TrbCtx* c, const double* clip12, const double* varyings, uint64_t n, const double* mv,
                              int kind, const void* uniforms, size_t ubytes) {
    NOT_WHILE_RECORDING(c, "submit_clip_triangles");
    if (!c || !c->in_frame || c->frame.nviews != 1) return fail(c, TRB_E_ARG, "submit: needs a single-view frame");
    if (!clip12 && n) return fail(c, TRB_E_ARG, "submit: null clip");
    if (n * 3 >= 0xFFFFFFF0ull || c->next_id + n >= 0xFFFFFFF0ull) return fail(c, TRB_E_ARG, "submit: too many triangles");
    const bool lit = kind == TRB_SHADER_PHONG || kind == TRB_SHADER_EYE;
    if (lit && !varyings) return fail(c, TRB_E_ARG, "submit: varyings required");
    int rc = check_device(c);
    if (rc) return rc;
    c->tris_submitted += n;
    if (n == 0) return TRB_OK;
    const void* dun = nullptr;
    if (kind == TRB_SHADER_SHADOW_PHONG || kind == TRB_SHADER_GOURAUD)
        return fail(c, TRB_E_SHADER, "submit: this shader needs a mesh draw");
    rc = resolve_uniforms(c, kind, uniforms, ubytes, 1, &dun);
    if (rc) return rc;
    cudaError_t e = cudaSuccess;
    const uint32_t nverts = (uint32_t)(n * 3);
    double* mats = (double*)c->arena.alloc(sizeof(double) * 32, e);
    CU(e);
    double hm[32];
    for (int i = 0; i < 32; ++i) hm[i] = (i % 16) % 5 == 0 ? 1.0 : 0.0;
    if (mv) memcpy(hm, mv, 128);
    CU(cudaMemcpyAsync(mats, hm, sizeof(hm), cudaMemcpyHostToDevice, c->stream));
    VRec* vrec = (VRec*)c->arena.alloc(sizeof(VRec) * (size_t)nverts, e);
    CU(e);
    double* dvary = nullptr;
    if (lit) {
        dvary = (double*)c->arena.alloc(sizeof(double) * 24 * n, e);
        CU(e);
        CU(cudaMemcpyAsync(dvary, varyings, sizeof(double) * 24 * n, cudaMemcpyHostToDevice, c->stream));
    }
    CU(c->scratch_a.ensure(sizeof(double) * 12 * n, c->stream));
    CU(cudaMemcpyAsync(c->scratch_a.p, clip12, sizeof(double) * 12 * n, cudaMemcpyHostToDevice, c->stream));
    {
        Launch L(c, "k_vertex_clip");
        k_vertex_clip<<<blocks_for(nverts), TPB, 0, c->stream>>>(c->frame, c->scratch_a.as<double>(), nverts, vrec);
    }
    CU(cudaGetLastError());
    GeomArgs g;
    g.idx = nullptr;
    g.first_tri = 0;
    g.ntris = (uint32_t)n;
    g.nverts = nverts;
    g.id_base = (uint32_t)c->next_id;
    g.vrec = vrec;
    g.perm = nullptr;
    g.idx_perm = nullptr;
    g.nslots = g.ntris;
    g.shard_n = g.shard_r = 0;
    g.shard_shift = 0;
    g.nperm = 0;
    rc = raster_draw(c, g);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));  // caller may reuse clip12 / varyings / hm
    DrawDev d{};
    d.id_base = g.id_base;
    d.ntris = g.ntris;
    d.first_tri = 0;
    d.nverts = nverts;
    d.idx = nullptr;
    d.attr8 = nullptr;
    d.vrec = vrec;
    d.mats = mats;
    d.uniforms = dun;
    d.varyings = dvary;
    d.kind = kind;
    d.mesh_ntris = g.ntris;
    d.mesh_id_base = (long long)g.id_base;
    c->draws.push_back(d);
    c->next_id += n;
    return TRB_OK;
}

int trb_depth_snapshot(TrbCtx* c) {
    HostSpan host_span_("trb_depth_snapshot");
    if (!c || !c->in_frame) return fail(c, TRB_E_ARG, "depth_snapshot: no frame");
    int rc = check_device(c);
    if (rc) return rc;
    size_t bytes = (size_t)c->frame.npix * c->frame.nviews * 8;
    CU(c->zsnap.ensure(bytes, c->stream));
    // unflushed pixels hold canonical keys; exact -0.0 bits only appear after their shade pass,
    // so resolve first to snapshot what the reference's vector copy (main.cpp:700) would hold
    rc = do_flush(c);
    if (rc) return rc;
    c->snap_lazy = c->lazy_snapshot;
    c->snap_all = false;
    if (c->snap_lazy) {
        const size_t nslots = (size_t)c->frame.nviews * c->frame.ntiles;
        CU(c->snap_saved.ensure(nslots, c->stream));
        Launch L(c, "clear_snapshot_tiles", nullptr, false);
        CU(cudaMemsetAsync(c->snap_saved.p, 0, nslots, c->stream));
    } else {
        Launch L(c, "copy_depth_snapshot", nullptr, false);
        CU(cudaMemcpyAsync(c->zsnap.p, c->zkey.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    c->have_snapshot = true;
    c->snap_stale = false;
    return TRB_OK;
}
int trb_depth_restore(TrbCtx* c) {
    HostSpan host_span_("trb_depth_restore");
    if (!c || !c->in_frame || !c->have_snapshot) return fail(c, TRB_E_ARG, "depth_restore: no snapshot");
    int rc = do_flush(c);  // colours of everything drawn so far persist (main.cpp:730)
    if (rc) return rc;
    if (c->snap_lazy) {                 // copy the saved tiles back; the others were never changed
        const FrameDev& f = c->frame;
        const uint32_t nslots = (uint32_t)((size_t)f.nviews * f.ntiles);
        Launch L(c, "k_snap_restore");
        k_snap_restore<<<(nslots + TPB - 1) / TPB, TPB, 0, c->stream>>>(f, nslots, c->snap_saved.as<uint8_t>(),
                                                                      c->zsnap.as<unsigned long long>());
        CU(cudaGetLastError());
        return TRB_OK;
    }
    if (c->snap_stale) return TRB_OK;   // restored already and nothing drawn since: the key plane IS the snapshot
    size_t bytes = (size_t)c->frame.npix * c->frame.nviews * 8;
    if (c->peers.n > 0) {               // peers hold the address of the key plane: copy
        CU(cudaMemcpyAsync(c->zkey.p, c->zsnap.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
        return TRB_OK;
    }
    // `zbuffer = zbuffer_before_eyes` (main.cpp:730) without moving 8 bytes per pixel: the saved plane
    // becomes the key plane; the old one is refreshed from it only if something is drawn before the next
    // snapshot / frame (refresh_snapshot), so that a later restore still finds the saved state
    std::swap(c->zkey, c->zsnap);
    c->frame.zkey = c->zkey.as<unsigned long long>();
    c->snap_stale = true;
    return TRB_OK;
}
int trb_keep_depth_as_shadow_map(TrbCtx* c, int32_t* out) {
    if (!c || !c->in_frame || !out) return fail(c, TRB_E_ARG, "keep_depth_as_shadow_map");
    int rc = do_flush(c);
    if (rc) return rc;
    ShadowMap b;
    size_t bytes = (size_t)c->frame.npix * 8;
    // planes of released maps are recycled (stream order keeps their last readers ahead of this copy):
    // a frame loop that builds a shadow map per frame neither allocates nor synchronises
    for (size_t i = 0; i < c->shadow_pool.size(); ++i)
        if (c->shadow_pool[i].cap >= bytes) {
            b.keys = c->shadow_pool[i];
            c->shadow_pool.erase(c->shadow_pool.begin() + i);
            break;
        }
    CU(b.keys.ensure(bytes, c->stream));
    CU(cudaMemcpyAsync(b.keys.p, c->zkey.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
    b.w = c->frame.W;
    b.h = c->frame.H;
    c->shadow_maps.push_back(b);
    *out = (int32_t)c->shadow_maps.size() - 1;
    return TRB_OK;
}

int trb_release_shadow_maps(TrbCtx* c) {
    NOT_WHILE_RECORDING(c, "release_shadow_maps");
    if (!c) return TRB_E_ARG;
    CU(cudaSetDevice(c->device));
    for (auto& b : c->shadow_maps) c->shadow_pool.push_back(b.keys);   // no sync, no cudaFree: reused by the next keep
    c->shadow_maps.clear();
    while (c->shadow_pool.size() > 4) {          // a caller that keeps many maps once should not pin them forever
        CU(cudaStreamSynchronize(c->stream));
        c->shadow_pool.back().release();
        c->shadow_pool.pop_back();
    }
    return TRB_OK;
}

int trb_flush(TrbCtx* c) {
    if (!c) return TRB_E_ARG;
    return do_flush(c);
}
int trb_end_frame(TrbCtx* c) {
    HostSpan host_span_("trb_end_frame");
    if (!c) return TRB_E_ARG;
    return do_flush(c);
}

// ---- frame recordings: a launch-bound frame as ONE CUDA graph launch ---------------------------------------------
// the frame a recording describes has (re)run: make the context look like it
static int recording_adopt_state(TrbCtx* c, Recording& r) {
    // shadow-map planes the frame kept go back from the pool (or stay) in the slots the frame gave them
    if (c->shadow_maps.size() < r.maps_before) return fail(c, TRB_E_ARG, "replay: shadow maps the recording uses were released");
    for (size_t i = 0; i < r.maps_kept.size(); ++i) {
        const size_t slot = r.maps_before + i;
        if (slot < c->shadow_maps.size()) {
            if (c->shadow_maps[slot].keys.p != r.maps_kept[i].keys.p)
                return fail(c, TRB_E_ARG, "replay: release the shadow maps of the frame before (trb_release_shadow_maps)");
            continue;
        }
        bool found = false;
        for (size_t k = 0; k < c->shadow_pool.size() && !found; ++k)
            if (c->shadow_pool[k].p == r.maps_kept[i].keys.p) {
                c->shadow_pool.erase(c->shadow_pool.begin() + k);
                found = true;
            }
        if (!found) return fail(c, TRB_E_ARG, "replay: a shadow-map plane of the recording is in use elsewhere");
        c->shadow_maps.push_back(r.maps_kept[i]);
    }
    c->frame = r.frame;
    c->zkey = r.zkey;
    c->zsnap = r.zsnap;
    c->have_snapshot = r.have_snapshot;
    c->snap_stale = r.snap_stale;
    c->snap_lazy = r.snap_lazy;
    c->snap_all = r.snap_all;
    c->foreign_ids = r.foreign_ids;
    c->next_id = r.next_id;
    c->tris_submitted = r.tris_submitted;
    c->shade_row0 = r.shade_row0;
    c->shade_row1 = r.shade_row1;
    c->draws.clear();
    c->in_frame = true;
    return TRB_OK;
}

int trb_record_begin(TrbCtx* c) {
    if (!c) return TRB_E_ARG;
    if (c->rec) return fail(c, TRB_E_ARG, "record_begin: already recording");
    if (c->sync_draws || c->profiling || c->comm.n > 0 || c->peers.n > 0)
        return fail(c, TRB_E_ARG, "record_begin: not with TRB_SYNC_DRAWS, kernel profiling or a composite group");
    int rc = check_device(c);
    if (rc) return rc;
    if (c->in_frame && (rc = do_flush(c))) return rc;      // what was drawn before is not part of the recording
    Recording* r = new Recording();
    r->params_cap = (size_t)4 << 20;
    if (cudaHostAlloc((void**)&r->params, r->params_cap, cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&r->done, cudaEventDisableTiming) != cudaSuccess) {
        recording_destroy(r);
        return fail(c, TRB_E_NOMEM, "record_begin: pinned parameter block");
    }
    r->maps_before = c->shadow_maps.size();
    cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) {
        recording_destroy(r);
        return fail(c, TRB_E_CUDA, std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e));
    }
    c->rec = r;
    c->launches_at_record = c->launches;
    ++t_recording;
    return TRB_OK;
}

int trb_record_end(TrbCtx* c, TrbRecording* out) {
    if (!c || !out) return fail(c, TRB_E_ARG, "record_end");
    if (!c->rec) return fail(c, TRB_E_ARG, "record_end: not recording");
    Recording* r = c->rec;
    *out = 0;
    int rc = check_device(c);
    if (!rc && r->failed.empty() && c->in_frame) rc = do_flush(c);   // a recording ends with its frame resolved
    if (!rc && !r->failed.empty()) {
        rc = TRB_E_ARG;
        c->err = "a call inside the recording failed: " + r->failed;
    }
    const std::string why = rc ? c->err : std::string();
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    c->rec = nullptr;
    if (t_recording > 0) --t_recording;
    if (!rc && e == cudaSuccess && g) e = cudaGraphInstantiate(&r->exec, g, 0);
    r->graph = g;
    if (rc || e != cudaSuccess || !r->exec) {
        recording_destroy(r);
        (void)cudaGetLastError();
        c->in_frame = false;                               // nothing of the frame ran: its calls were only captured
        c->draws.clear();
        return rc ? fail(c, rc, "record_end: " + why)
                  : fail(c, TRB_E_CUDA, std::string("record_end: the capture failed: ") + cudaGetErrorString(e));
    }
    r->launches = c->launches - c->launches_at_record;
    r->frame = c->frame;
    r->zkey = c->zkey;
    r->zsnap = c->zsnap;
    r->have_snapshot = c->have_snapshot;
    r->snap_stale = c->snap_stale;
    r->snap_lazy = c->snap_lazy;
    r->snap_all = c->snap_all;
    r->foreign_ids = c->foreign_ids;
    r->next_id = c->next_id;
    r->tris_submitted = c->tris_submitted;
    r->shade_row0 = c->shade_row0;
    r->shade_row1 = c->shade_row1;
    for (size_t i = r->maps_before; i < c->shadow_maps.size(); ++i) r->maps_kept.push_back(c->shadow_maps[i]);
    r->generation = g_generation.load();
    // capturing ran nothing: run the frame now, so that the context is where the calls left it
    CU(cudaGraphLaunch(r->exec, c->stream));
    CU(cudaEventRecord(r->done, c->stream));
    size_t slot = 0;
    while (slot < c->recordings.size() && c->recordings[slot]) ++slot;
    if (slot == c->recordings.size()) c->recordings.push_back(r); else c->recordings[slot] = r;
    *out = slot + 1;
    return TRB_OK;
}

int trb_replay(TrbCtx* c, TrbRecording h, const TrbReplayDraw* draws, int ndraws) {
    HostSpan host_span_("trb_replay");
    if (!c || h == 0 || h > c->recordings.size() || !c->recordings[h - 1]) return fail(c, TRB_E_ARG, "replay: bad recording");
    NOT_WHILE_RECORDING(c, "replay");
    Recording& r = *c->recordings[h - 1];
    if (r.generation != g_generation.load())
        return fail(c, TRB_E_ARG, "replay: the recording is stale (a buffer, mesh or texture it refers to was reallocated or freed)");
    int rc = check_device(c);
    if (rc) return rc;
    if (c->in_frame && (rc = do_flush(c))) return rc;      // the frame in flight resolves first
    rc = recording_adopt_state(c, r);                      // first: the uniforms below are validated against this state
    if (rc) return rc;
    if (draws) {
        if (ndraws != (int)r.draws.size()) return fail(c, TRB_E_ARG, "replay: one entry per draw call of the recording");
        CU(cudaEventSynchronize(r.done));                  // the replay before this one has read the parameter block
        std::vector<char> ub;
        std::vector<trbf::LitF> hl;
        for (int i = 0; i < ndraws; ++i) {
            Recording::Draw& s = r.draws[i];
            const TrbReplayDraw& d = draws[i];
            if (s.mats_off == (size_t)-1) continue;        // an empty draw
            double* mats = reinterpret_cast<double*>(r.params + s.mats_off);
            for (int v = 0; v < s.nviews; ++v) {
                if (d.modelview) memcpy(mats + (size_t)v * 32, d.modelview + 16 * v, 128);
                if (d.perspective) memcpy(mats + (size_t)v * 32 + 16, d.perspective + 16 * v, 128);
            }
            if (d.modelview) s.mv.assign(d.modelview, d.modelview + (size_t)16 * s.nviews);
            if (d.uniforms) {
                if (d.uniform_bytes != s.uni_host_bytes || s.uni_off == (size_t)-1)
                    return fail(c, TRB_E_ARG, "replay: uniform block size differs from the recorded draw's");
                rc = build_uniforms(c, s.kind, d.uniforms, d.uniform_bytes, s.nviews, ub);
                if (rc) return rc;
                memcpy(r.params + s.uni_off, ub.data(), ub.size());
                s.uni.assign((const char*)d.uniforms, (const char*)d.uniforms + d.uniform_bytes * s.nviews);
            }
            if (s.litf_off != (size_t)-1 && (d.modelview || d.uniforms)) {
                build_litf(c, s.kind, s.uni.data(), s.mv.data(), s.nviews, hl);
                memcpy(r.params + s.litf_off, hl.data(), hl.size() * sizeof(trbf::LitF));
            }
        }
    }
    CU(cudaGraphLaunch(r.exec, c->stream));
    CU(cudaEventRecord(r.done, c->stream));
    c->launches += r.launches;
    return TRB_OK;
}

int trb_recording_free(TrbCtx* c, TrbRecording h) {
    if (!c || h == 0 || h > c->recordings.size() || !c->recordings[h - 1]) return fail(c, TRB_E_ARG, "recording_free");
    NOT_WHILE_RECORDING(c, "recording_free");
    int rc = check_device(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    recording_destroy(c->recordings[h - 1]);
    c->recordings[h - 1] = nullptr;
    return TRB_OK;
}

int trb_read_color(TrbCtx* c, int view, uint8_t* out) {
    NOT_WHILE_RECORDING(c, "read_color");
    if (!c || !c->in_frame || view < 0 || view >= c->frame.nviews || !out) return fail(c, TRB_E_ARG, "read_color");
    int rc = do_flush(c);
    if (rc) return rc;
    size_t bytes = (size_t)c->frame.npix * 3;
    CU(cudaMemcpyAsync(out, c->color.as<uint8_t>() + bytes * view, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TRB_OK;
}
int trb_write_color(TrbCtx* c, int view, const uint8_t* bgr) {
    NOT_WHILE_RECORDING(c, "write_color");
    if (!c || !c->in_frame || view < 0 || view >= c->frame.nviews || !bgr) return fail(c, TRB_E_ARG, "write_color");
    int rc = do_flush(c);
    if (rc) return rc;
    size_t bytes = (size_t)c->frame.npix * 3;
    CU(cudaMemcpyAsync(c->color.as<uint8_t>() + bytes * view, bgr, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));   // the caller's image may be pageable and reused at once
    return TRB_OK;
}
int trb_read_depth(TrbCtx* c, int view, double* out) {
    NOT_WHILE_RECORDING(c, "read_depth");
    if (!c || !c->in_frame || view < 0 || view >= c->frame.nviews || !out) return fail(c, TRB_E_ARG, "read_depth");
    int rc = do_flush(c);
    if (rc) return rc;
    const unsigned long long n = c->frame.npix;
    CU(c->scratch_b.ensure(n * 8, c->stream));
    {
        Launch L(c, "k_unmap_depth");
        k_unmap_depth<<<blocks_for(n), TPB, 0, c->stream>>>(c->frame.zkey + n * view, n, c->scratch_b.as<double>());
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->scratch_b.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TRB_OK;
}
int trb_readback_wait(TrbCtx* c) {
    NOT_WHILE_RECORDING(c, "readback_wait");
    if (!c) return TRB_E_ARG;
    CU(cudaSetDevice(c->device));
    for (int i = 0; i < TrbCtx::RB_SLOTS; ++i)
        if (c->rb_inflight[i]) {
            CU(cudaEventSynchronize(c->rb_done[i]));
            c->rb_inflight[i] = false;
        }
    for (int i = 0; i < 2; ++i) {
        int rc = tga_finish(c, c->tga[i], /*wait_copies=*/true);
        if (rc) return rc;
    }
    return TRB_OK;
}
int trb_readback_async(TrbCtx* c, uint8_t* const* color_out, double* const* depth_out) {
    NOT_WHILE_RECORDING(c, "readback_async");
    HostSpan host_span_("trb_readback_async");
    if (!c || !c->in_frame) return fail(c, TRB_E_ARG, "readback_async: no frame");
    int rc = do_flush(c);
    if (rc) return rc;
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < TrbCtx::RB_SLOTS; ++i) {
            CU(cudaEventCreateWithFlags(&c->rb_ready[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->rb_done[i], cudaEventDisableTiming));
        }
    }
    const int i = c->rb_idx = (c->rb_idx + 1) % TrbCtx::RB_SLOTS;
    if (c->rb_inflight[i]) {  // staging area i is still being drained by the copy stream
        HostSpan w("readback_async:wait_slot");
        CU(cudaEventSynchronize(c->rb_done[i]));
        c->rb_inflight[i] = false;
    }
    const FrameDev& f = c->frame;
    const size_t total = (size_t)f.npix * f.nviews;
    const size_t color_bytes = color_out ? total * 3 : 0, depth_off = (color_bytes + 255) & ~(size_t)255;
    CU(c->rb[i].ensure(depth_off + (depth_out ? total * 8 : 0), c->stream));
    uint8_t* stage = c->rb[i].as<uint8_t>();
    if (color_out) {
        Launch L(c, "copy_stage_color", nullptr, false);
        CU(cudaMemcpyAsync(stage, f.color, color_bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (depth_out) {
        Launch L(c, "k_unmap_depth");
        k_unmap_depth<<<blocks_for(total), TPB, 0, c->stream>>>(f.zkey, total, reinterpret_cast<double*>(stage + depth_off));
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->rb_ready[i], c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->rb_ready[i], 0));
    Launch span(c, "copy_d2h_frames", c->copy_stream, false);
    HostSpan d2h("readback_async:enqueue_d2h");
    for (int v = 0; v < f.nviews; ++v) {
        if (color_out && color_out[v])
            CU(cudaMemcpyAsync(color_out[v], stage + (size_t)f.npix * 3 * v, (size_t)f.npix * 3, cudaMemcpyDeviceToHost,
                               c->copy_stream));
        if (depth_out && depth_out[v])
            CU(cudaMemcpyAsync(depth_out[v], stage + depth_off + (size_t)f.npix * 8 * v, (size_t)f.npix * 8,
                               cudaMemcpyDeviceToHost, c->copy_stream));
    }
    CU(cudaEventRecord(c->rb_done[i], c->copy_stream));
    c->rb_inflight[i] = true;
    return TRB_OK;
}
int trb_read_visibility(TrbCtx* c, int view, uint32_t* out) {
    NOT_WHILE_RECORDING(c, "read_visibility");
    if (!c || !c->in_frame || view < 0 || view >= c->frame.nviews || !out) return fail(c, TRB_E_ARG, "read_visibility");
    int rc = check_device(c);
    if (rc) return rc;
    size_t n = (size_t)c->frame.npix;
    CU(cudaMemcpyAsync(out, c->frame.vis + n * view, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TRB_OK;
}

int trb_get_stats(TrbCtx* c, int view, TrbStats* out) {
    NOT_WHILE_RECORDING(c, "get_stats");
    if (!c || !c->in_frame || view < 0 || view >= c->frame.nviews || !out) return fail(c, TRB_E_ARG, "get_stats");
    int rc = check_device(c);
    if (rc) return rc;
    const unsigned long long n = c->frame.npix;
    CU(c->scratch_b.ensure(64, c->stream));
    unsigned long long init[3] = {0ull, ~0ull, 0ull};  // finite count, min key, max key
    CU(cudaMemcpyAsync(c->scratch_b.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    {
        unsigned grid = (unsigned)std::min<unsigned long long>(blocks_for(n), (unsigned long long)c->sms * 16);
        Launch L(c, "k_count_finite");
        k_count_finite<<<grid, TPB, 0, c->stream>>>(c->frame.zkey + n * view, n, c->scratch_b.as<unsigned long long>());
    }
    {
        unsigned grid = (unsigned)std::min<unsigned long long>(blocks_for(n), (unsigned long long)c->sms * 16);
        Launch L(c, "k_depth_range");
        k_depth_range<<<grid, TPB, 0, c->stream>>>(c->frame.zkey + n * view, n, c->scratch_b.as<unsigned long long>() + 1);
    }
    CU(cudaGetLastError());
    DevStats s;
    unsigned long long res[3] = {0, 0, 0};
    CU(cudaMemcpyAsync(&s, c->frame.stats + view, sizeof(s), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(res, c->scratch_b.p, sizeof(res), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const unsigned long long finite = res[0];
    memset(out, 0, sizeof(*out));
    out->triangles_submitted = c->tris_submitted;
    out->triangles_binned = s.tri_binned;
    out->tile_entries = s.tile_entries;
    out->fragments_covered = s.frag_covered;
    out->pixels_shaded = finite;
    out->bbox_min_x = s.bx0;
    out->bbox_min_y = s.by0;
    out->bbox_max_x = s.bx1;
    out->bbox_max_y = s.by1;
    // z_min is tracked by the raster kernel over drawn fragments (it must survive a depth_restore,
    // like the reference's static min_z); the max is the largest depth left in the buffer
    out->z_min = s.zmin_key != ~0ull ? depth_from_key(s.zmin_key) : INFINITY;
    out->z_max_covered = finite ? depth_from_key(res[2]) : -INFINITY;
    out->z_max_ref = NAN;
    return TRB_OK;
}

int trb_synchronize(TrbCtx* c) {
    NOT_WHILE_RECORDING(c, "synchronize");
    if (!c) return TRB_E_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return TRB_OK;
}

// ---- post passes -----------------------------------------------------------------------------
}  // extern "C"

namespace {
// one post-pass image of `view` into device memory: TRB_IMAGE_SSAO / _DEPTH (1 byte per pixel),
// TRB_IMAGE_FINAL (3 bytes per pixel).  The frame must be flushed.
int post_plane(TrbCtx* c, int which, int view, uint8_t* dst) {
    const FrameDev& f = c->frame;
    SsaoDirs d;
    ssao_dirs(d);
    if (which == TRB_IMAGE_SSAO) {
        Launch L(c, "k_ssao");
        k_ssao<<<dim3(f.tw, f.th), TPB, 0, c->stream>>>(f.zkey + f.npix * view, f.W, f.H, d, dst);
    } else if (which == TRB_IMAGE_FINAL) {
        Launch L(c, "k_composite_ao");
        k_composite_ao<<<dim3(f.tw, f.th), TPB, 0, c->stream>>>(f.zkey + f.npix * view, f.color + f.npix * 3 * view,
                                                                 f.W, f.H, d, dst);
    } else if (which == TRB_IMAGE_DEPTH) {
        CU(c->scratch_b.ensure(64, c->stream));
        unsigned long long init[2] = {~0ull, 0ull};
        CU(cudaMemcpyAsync(c->scratch_b.p, init, 16, cudaMemcpyHostToDevice, c->stream));
        {
            unsigned grid = (unsigned)std::min<unsigned long long>(blocks_for(f.npix), (unsigned long long)c->sms * 16);
            Launch L(c, "k_depth_range");
            k_depth_range<<<grid, TPB, 0, c->stream>>>(f.zkey + f.npix * view, f.npix, c->scratch_b.as<unsigned long long>());
        }
        {
            Launch L(c, "k_depth_image");
            k_depth_image<<<blocks_for(f.npix), TPB, 0, c->stream>>>(f.zkey + f.npix * view, f.npix,
                                                                      c->scratch_b.as<unsigned long long>(), dst);
        }
    } else {
        return fail(c, TRB_E_ARG, "post pass: unknown image");
    }
    CU(cudaGetLastError());
    return TRB_OK;
}
int post_to_host(TrbCtx* c, int which, int view, uint8_t* out, const char* what) {
    NOT_WHILE_RECORDING(c, "post passes");
    if (!c || !c->in_frame || view < 0 || view >= c->frame.nviews || !out) return fail(c, TRB_E_ARG, what);
    int rc = do_flush(c);
    if (rc) return rc;
    const size_t bytes = (size_t)c->frame.npix * (which == TRB_IMAGE_FINAL ? 3 : 1);
    CU(c->scratch_a.ensure(bytes, c->stream));
    rc = post_plane(c, which, view, c->scratch_a.as<uint8_t>());
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, c->scratch_a.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TRB_OK;
}

// the five passes of tga_rle.cuh over `total` = nviews * npix pixels of BPP bytes at `px`
template <int BPP>
int rle_passes(TrbCtx* c, const uint8_t* px, size_t npix, uint32_t nviews, uint8_t* out, uint32_t* view_offsets_dev) {
    using namespace trbr;
    const size_t total = npix * nviews;
    const size_t ns_max = total / 2 + nviews + 2;                     // a long run is at least two pixels
    // one allocation: flags | ls_excl | seg_start | long_end | bytes | base | map | prefix | scan aggregates | totals
    const size_t nb_px = (total + RBLOCK - 1) / RBLOCK, nb_seg = (ns_max + RBLOCK - 1) / RBLOCK;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_ls = take(total * 4), o_start = take(ns_max * 4), o_end = take(ns_max * 4), o_bytes = take(ns_max * 4),
                 o_base = take(ns_max * 4), o_agg = take(std::max(nb_px, nb_seg) * 4), o_tot = take(64),
                 o_flags = take(total), o_map = take(ns_max), o_prefix = take(ns_max);
    CU(c->rle_work.ensure(off, c->stream));
    char* w = c->rle_work.as<char>();
    uint8_t* flags = (uint8_t*)(w + o_flags);
    uint32_t* ls_excl = (uint32_t*)(w + o_ls);
    uint32_t* agg = (uint32_t*)(w + o_agg);
    uint32_t* ls_total = (uint32_t*)(w + o_tot);
    RleTables t;
    t.seg_start = (uint32_t*)(w + o_start);
    t.long_end = (uint32_t*)(w + o_end);
    t.bytes = (uint32_t*)(w + o_bytes);
    t.base = (uint32_t*)(w + o_base);
    t.map = (uint8_t*)(w + o_map);
    t.prefix = (uint8_t*)(w + o_prefix);
    const unsigned gpx = (unsigned)((total + RTPB - 1) / RTPB), gseg = (unsigned)((ns_max + RTPB - 1) / RTPB);
    // segments past the real count (only known on the device) are given neutral inputs: a zero map would
    // be wrong, so their tables are zeroed and the per-segment kernels stop at the device-side count
    CU(cudaMemsetAsync(t.bytes, 0, ns_max * 4, c->stream));
    CU(cudaMemsetAsync(t.map, 2, ns_max, c->stream));                 // identity map
    {
        Launch L(c, "k_rle_flags");
        k_rle_flags<BPP><<<gpx, RTPB, 0, c->stream>>>(px, npix, total, flags);
    }
    {   // 2: segment index of every pixel
        Launch L(c, "k_rle_scan");
        LoadLongStart ld{flags};
        k_rscan_partial<AddU32, LoadLongStart><<<(unsigned)nb_px, RTPB, 0, c->stream>>>(ld, total, agg);
        k_rscan_sums<AddU32><<<1, RTPB, 0, c->stream>>>(agg, (uint32_t)nb_px, ls_total);
        k_rscan_final<AddU32, LoadLongStart><<<(unsigned)nb_px, RTPB, 0, c->stream>>>(ld, total, agg, ls_excl);
    }
    {
        Launch L(c, "k_rle_segments");
        k_rle_segments<<<gpx, RTPB, 0, c->stream>>>(flags, ls_excl, ls_total, npix, total, nviews, t);
    }
    {   // 3: entry class of every segment
        Launch L(c, "k_rle_maps");
        k_rle_maps<<<gseg, RTPB, 0, c->stream>>>(ls_total, nviews, t);
        LoadPlain<uint8_t> ld{t.map};
        uint8_t* agg8 = (uint8_t*)agg;
        k_rscan_partial<MapCompose, LoadPlain<uint8_t>><<<(unsigned)nb_seg, RTPB, 0, c->stream>>>(ld, ns_max, agg8);
        k_rscan_sums<MapCompose><<<1, RTPB, 0, c->stream>>>(agg8, (uint32_t)nb_seg, (uint8_t*)(ls_total + 2));
        k_rscan_final<MapCompose, LoadPlain<uint8_t>><<<(unsigned)nb_seg, RTPB, 0, c->stream>>>(ld, ns_max, agg8, t.prefix);
    }
    {   // 4: where every segment writes
        Launch L(c, "k_rle_sizes");
        k_rle_sizes<BPP><<<gseg, RTPB, 0, c->stream>>>(ls_total, nviews, t);
        LoadPlain<uint32_t> ld{t.bytes};
        k_rscan_partial<AddU32, LoadPlain<uint32_t>><<<(unsigned)nb_seg, RTPB, 0, c->stream>>>(ld, ns_max, agg);
        k_rscan_sums<AddU32><<<1, RTPB, 0, c->stream>>>(agg, (uint32_t)nb_seg, ls_total + 1);
        k_rscan_final<AddU32, LoadPlain<uint32_t>><<<(unsigned)nb_seg, RTPB, 0, c->stream>>>(ld, ns_max, agg, t.base);
    }
    {
        Launch L(c, "k_rle_emit");
        k_rle_emit<BPP><<<gpx, RTPB, 0, c->stream>>>(px, flags, ls_excl, npix, total, t, out);
        k_rle_view_offsets<<<(nviews + 1 + 255) / 256, 256, 0, c->stream>>>(ls_excl, ls_total, npix, nviews, t, view_offsets_dev);
    }
    CU(cudaGetLastError());
    return TRB_OK;
}
}  // namespace

extern "C" {

int trb_ssao(TrbCtx* c, int view, uint8_t* out) { return post_to_host(c, TRB_IMAGE_SSAO, view, out, "ssao"); }
int trb_composite_ao(TrbCtx* c, int view, uint8_t* out) { return post_to_host(c, TRB_IMAGE_FINAL, view, out, "composite_ao"); }
int trb_depth_image(TrbCtx* c, int view, uint8_t* out) { return post_to_host(c, TRB_IMAGE_DEPTH, view, out, "depth_image"); }

}  // extern "C"

namespace {
// flush, build the source image of `which` for every view and packetise it into `outbuf` on the render stream;
// *offs_dev = nviews + 1 byte offsets of the views' packet runs inside outbuf (device memory, behind the packets)
int encode_on_device(TrbCtx* c, int which, DevBuf& outbuf, uint32_t** offs_dev) {
    int rc = do_flush(c);
    if (rc) return rc;
    const FrameDev& f = c->frame;
    const int bpp = 3;   // all four are TGAImage::RGB files: the grey maps store TGAColor(v, v, v) (main.cpp:309, 761)
    const size_t total = (size_t)f.npix * f.nviews;
    if (total * (bpp + 1) >= 0xFFFFFFF0ull) return fail(c, TRB_E_ARG, "encode_tga: batch too large for 32-bit offsets");
    if (f.W > 65535 || f.H > 65535) return fail(c, TRB_E_ARG, "encode_tga: TGA dimensions are 16 bit");
    const uint8_t* src = f.color;
    if (which != TRB_IMAGE_COLOR) {
        CU(c->rle_src.ensure(total * bpp, c->stream));
        const bool grey = which != TRB_IMAGE_FINAL;
        if (grey) CU(c->scratch_a.ensure(f.npix, c->stream));
        for (int v = 0; v < f.nviews; ++v) {
            uint8_t* dst = c->rle_src.as<uint8_t>() + (size_t)f.npix * bpp * v;
            rc = post_plane(c, which, v, grey ? c->scratch_a.as<uint8_t>() : dst);
            if (rc) return rc;
            if (grey) {
                Launch L(c, "k_grey_to_bgr");
                k_grey_to_bgr<<<blocks_for(f.npix), TPB, 0, c->stream>>>(c->scratch_a.as<uint8_t>(), f.npix, dst);
            }
        }
        CU(cudaGetLastError());
        src = c->rle_src.as<uint8_t>();
    }
    const size_t out_cap = total * bpp + total / 2 + f.nviews + 1024;  // worst case: a raw packet of two pixels per header
    const size_t offs_at = (out_cap + 3) & ~(size_t)3;                // per-view byte ranges behind the packets
    CU(outbuf.ensure(offs_at + ((size_t)f.nviews + 2) * 4, c->stream));
    *offs_dev = reinterpret_cast<uint32_t*>(outbuf.as<uint8_t>() + offs_at);
    return rle_passes<3>(c, src, f.npix, (uint32_t)f.nviews, outbuf.as<uint8_t>(), *offs_dev);
}
// TGAHeader of write_tga_file (tgaimage.cpp:167-178): 18 packed bytes, origin bottom-left (vflip = true)
void tga_header(uint8_t* h, int W, int H) {
    memset(h, 0, 18);
    h[2] = 10;                                                    // datatypecode: RLE true-colour
    h[12] = (uint8_t)(W & 255); h[13] = (uint8_t)(W >> 8);
    h[14] = (uint8_t)(H & 255); h[15] = (uint8_t)(H >> 8);
    h[16] = 24;
    h[17] = 0x00;
}
// the copies of an asynchronous encode whose per-view sizes have reached the host: headers + sizes now, packets on the
// copy stream
int tga_issue_copies(TrbCtx* c, TrbCtx::TgaJob& j) {
    j.pending_copy = false;
    for (int v = 0; v < j.nviews; ++v) {
        const uint64_t n = (uint64_t)j.offs_host[v + 1] - j.offs_host[v];
        j.sizes[v] = 18 + n;
        if (!j.outs[v]) continue;
        if (j.sizes[v] > j.capacity) return fail(c, TRB_E_ARG, "encode_tga: output buffer too small");
        tga_header(j.outs[v], j.W, j.H);
        CU(cudaMemcpyAsync(j.outs[v] + 18, j.out.as<uint8_t>() + j.offs_host[v], n, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    CU(cudaEventRecord(j.done, c->copy_stream));
    return TRB_OK;
}
int tga_finish(TrbCtx* c, TrbCtx::TgaJob& j, bool wait_copies) {
    if (!j.inflight) return TRB_OK;
    if (j.pending_copy) {
        CU(cudaEventSynchronize(j.offs_ready));
        int rc = tga_issue_copies(c, j);
        if (rc) return rc;
    }
    if (wait_copies) {
        CU(cudaEventSynchronize(j.done));
        j.inflight = false;
    }
    return TRB_OK;
}
}  // namespace

extern "C" {

int trb_encode_tga(TrbCtx* c, int which, uint8_t* const* out, uint64_t capacity, uint64_t* sizes) {
    NOT_WHILE_RECORDING(c, "encode_tga");
    if (!c || !c->in_frame || !out || !sizes || which < TRB_IMAGE_COLOR || which > TRB_IMAGE_FINAL)
        return fail(c, TRB_E_ARG, "encode_tga: bad argument");
    uint32_t* offs_dev = nullptr;
    int rc = encode_on_device(c, which, c->rle_out, &offs_dev);
    if (rc) return rc;
    const FrameDev& f = c->frame;
    std::vector<uint32_t> offs((size_t)f.nviews + 1);
    CU(cudaMemcpyAsync(offs.data(), offs_dev, offs.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int v = 0; v < f.nviews; ++v) {
        const uint64_t n = (uint64_t)offs[v + 1] - offs[v];
        sizes[v] = 18 + n;
        if (!out[v]) continue;
        if (sizes[v] > capacity) return fail(c, TRB_E_ARG, "encode_tga: output buffer too small");
        tga_header(out[v], f.W, f.H);
        CU(cudaMemcpyAsync(out[v] + 18, c->rle_out.as<uint8_t>() + offs[v], n, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return TRB_OK;
}

// The asynchronous frame writer: packets are built behind the frame on the render stream, their per-view sizes go to
// pinned host memory, and the packets themselves come home on the copy stream as soon as the host knows the sizes -
// at the latest inside the next trb_encode_tga_async / trb_readback_wait.  Two encodes may be in flight.
int trb_encode_tga_async(TrbCtx* c, int which, uint8_t* const* out, uint64_t capacity, uint64_t* sizes) {
    NOT_WHILE_RECORDING(c, "encode_tga_async");
    HostSpan host_span_("trb_encode_tga_async");
    if (!c || !c->in_frame || !out || !sizes || which < TRB_IMAGE_COLOR || which > TRB_IMAGE_FINAL)
        return fail(c, TRB_E_ARG, "encode_tga_async: bad argument");
    int rc = check_device(c);
    if (rc) return rc;
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < TrbCtx::RB_SLOTS; ++i) {
            CU(cudaEventCreateWithFlags(&c->rb_ready[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->rb_done[i], cudaEventDisableTiming));
        }
    }
    const int i = (c->tga_idx ^= 1);
    TrbCtx::TgaJob& j = c->tga[i];
    TrbCtx::TgaJob& other = c->tga[i ^ 1];
    // the previous encode's sizes have usually arrived by now: send its packets on their way first
    if (other.inflight && other.pending_copy && cudaEventQuery(other.offs_ready) == cudaSuccess) {
        rc = tga_issue_copies(c, other);
        if (rc) return rc;
    }
    (void)cudaGetLastError();
    rc = tga_finish(c, j, /*wait_copies=*/true);      // slot i was used two encodes ago: its packets must be home
    if (rc) return rc;
    const FrameDev& f = c->frame;
    if (!j.offs_ready) {
        CU(cudaEventCreateWithFlags(&j.offs_ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&j.done, cudaEventDisableTiming));
    }
    if (j.offs_cap < (size_t)f.nviews + 1) {
        if (j.offs_host) cudaFreeHost(j.offs_host);
        CU(cudaHostAlloc((void**)&j.offs_host, ((size_t)f.nviews + 1) * 4, cudaHostAllocDefault));
        j.offs_cap = (size_t)f.nviews + 1;
    }
    uint32_t* offs_dev = nullptr;
    rc = encode_on_device(c, which, j.out, &offs_dev);
    if (rc) return rc;
    CU(cudaMemcpyAsync(j.offs_host, offs_dev, ((size_t)f.nviews + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(j.offs_ready, c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, j.offs_ready, 0));   // the copy stream must not read the packets earlier
    j.outs.assign(out, out + f.nviews);
    j.sizes = sizes;
    j.capacity = capacity;
    j.nviews = f.nviews;
    j.W = f.W;
    j.H = f.H;
    j.pending_copy = true;
    j.inflight = true;
    return TRB_OK;
}

// ---- timing -------------------------------------------------------------------------------------
int trb_timer_start(TrbCtx* c) {
    NOT_WHILE_RECORDING(c, "timer_start");
    if (!c) return TRB_E_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev_a, c->stream));
    return TRB_OK;
}
int trb_timer_stop_ms(TrbCtx* c, float* ms) {
    NOT_WHILE_RECORDING(c, "timer_stop_ms");
    if (!c || !ms) return TRB_E_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev_b, c->stream));
    CU(cudaEventSynchronize(c->ev_b));
    CU(cudaEventElapsedTime(ms, c->ev_a, c->ev_b));
    return TRB_OK;
}
int trb_profile_enable(TrbCtx* c, int on) {
    NOT_WHILE_RECORDING(c, "profile_enable");
    if (!c) return TRB_E_ARG;
    cudaSetDevice(c->device);
    prof_collect(c);
    c->profiling = on != 0;
    return TRB_OK;
}
int trb_profile_read(TrbCtx* c, TrbKernelTime* out, int capacity, int* n_out, int reset) {
    NOT_WHILE_RECORDING(c, "profile_read");
    if (!c || !n_out) return TRB_E_ARG;
    cudaSetDevice(c->device);
    prof_collect(c);
    int n = 0;
    for (auto& a : c->prof_acc) {
        if (n >= capacity || !out) break;
        memset(&out[n], 0, sizeof(out[n]));
        strncpy(out[n].name, a.name.c_str(), sizeof(out[n].name) - 1);
        out[n].launches = a.launches;
        out[n].ms = a.ms;
        ++n;
    }
    *n_out = n;
    if (reset) c->prof_acc.clear();
    return TRB_OK;
}
uint64_t trb_launch_count(TrbCtx* c) { return c ? c->launches : 0; }

// ---- multi-GPU composite ----------------------------------------------------------------------
int trb_device_planes(TrbCtx* c, uint64_t* key_ptr, uint64_t* vis_ptr, uint64_t* npix) {
    if (!c || !c->in_frame || !key_ptr || !vis_ptr || !npix) return fail(c, TRB_E_COMM, "device_planes: no frame");
    *key_ptr = (uint64_t)(uintptr_t)c->frame.zkey;
    *vis_ptr = (uint64_t)(uintptr_t)c->frame.vis;
    *npix = c->frame.npix;
    return TRB_OK;
}
int trb_set_triangle_id_base(TrbCtx* c, uint64_t base) {
    NOT_WHILE_RECORDING(c, "set_triangle_id_base");
    if (!c || base >= 0xFFFFFFF0ull) return fail(c, TRB_E_ARG, "set_triangle_id_base");
    c->next_id = base;
    c->foreign_ids = true;
    return TRB_OK;
}
// The host all-reduces (min) the depth-key plane in place as int64; keys are made signed-sortable
// for the collective and restored by composite_mask.
int trb_composite_save_local_depth(TrbCtx* c) {
    NOT_WHILE_RECORDING(c, "composite_save_local_depth");
    if (!c || !c->in_frame || c->frame.nviews != 1) return fail(c, TRB_E_COMM, "composite: needs a single-view frame");
    for (const DrawDev& d : c->draws)
        if (d.vmark) return fail(c, TRB_E_COMM, "composite: shares drawn while a composite group is open are composited with trb_composite "
                                               "(their vertex stage only covers the share)");
    int rc = check_device(c);
    if (rc) return rc;
    if (snapshot_window(c) && (rc = snapshot_save_tiles(c, nullptr, nullptr))) return rc;   // the key plane is rewritten below
    const unsigned long long n = c->frame.npix;
    CU(c->zlocal.ensure(n * 8, c->stream));
    CU(cudaMemcpyAsync(c->zlocal.p, c->zkey.p, n * 8, cudaMemcpyDeviceToDevice, c->stream));
    {
        Launch L(c, "k_key_to_sortable_i64");
        k_key_to_sortable_i64<<<blocks_for(n), TPB, 0, c->stream>>>(c->frame.zkey, n);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return TRB_OK;
}
int trb_composite_mask(TrbCtx* c) {
    NOT_WHILE_RECORDING(c, "composite_mask");
    if (!c || !c->in_frame || c->frame.nviews != 1 || c->zlocal.cap < c->frame.npix * 8)
        return fail(c, TRB_E_COMM, "composite_mask: call composite_save_local_depth first");
    int rc = check_device(c);
    if (rc) return rc;
    const unsigned long long n = c->frame.npix;
    {
        Launch L(c, "k_key_to_sortable_i64");
        k_key_to_sortable_i64<<<blocks_for(n), TPB, 0, c->stream>>>(c->frame.zkey, n);
    }
    {
        Launch L(c, "k_composite_mask");
        k_composite_mask<<<blocks_for(n), TPB, 0, c->stream>>>(c->zlocal.as<unsigned long long>(), c->frame.zkey,
                                                               c->frame.vis, n);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return TRB_OK;
}
int trb_composite_finish(TrbCtx* c) {
    NOT_WHILE_RECORDING(c, "composite_finish");
    if (!c || !c->in_frame || c->frame.nviews != 1) return fail(c, TRB_E_COMM, "composite_finish: needs a single-view frame");
    int rc = check_device(c);
    if (rc) return rc;
    const unsigned long long n = c->frame.npix;
    {
        Launch L(c, "k_composite_finish");
        k_composite_finish<<<blocks_for(n), TPB, 0, c->stream>>>(c->frame.vis, n);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return TRB_OK;
}
int trb_ipc_export_planes(TrbCtx* c, void* key_handle, void* vis_handle) {
    NOT_WHILE_RECORDING(c, "ipc_export_planes");
    if (!c || !c->in_frame || !key_handle || !vis_handle) return fail(c, TRB_E_COMM, "ipc_export_planes: no frame");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    int rc = check_device(c);
    if (rc) return rc;
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)key_handle, c->zkey.p));
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)vis_handle, c->vis.p));
    return TRB_OK;
}
int trb_ipc_close_peers(TrbCtx* c) {
    if (!c) return TRB_E_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (void*& p : c->peer_open)
        if (p) {
            cudaIpcCloseMemHandle(p);
            p = nullptr;
        }
    c->peers.n = 0;
    c->peer_rank = -1;
    return TRB_OK;
}
int trb_ipc_open_peers(TrbCtx* c, const void* key_handles, const void* vis_handles, int n, int my_rank) {
    NOT_WHILE_RECORDING(c, "ipc_open_peers");
    if (!c || !c->in_frame || !key_handles || !vis_handles || n < 1 || n > MAX_PEERS || my_rank < 0 || my_rank >= n)
        return fail(c, TRB_E_COMM, "ipc_open_peers: bad argument");
    int rc = check_device(c);
    if (rc) return rc;
    trb_ipc_close_peers(c);
    const cudaIpcMemHandle_t* kh = (const cudaIpcMemHandle_t*)key_handles;
    const cudaIpcMemHandle_t* vh = (const cudaIpcMemHandle_t*)vis_handles;
    for (int r = 0; r < n; ++r) {
        c->peers.touched[r] = nullptr;
        if (r == my_rank) {
            c->peers.key[r] = c->frame.zkey;
            c->peers.vis[r] = c->frame.vis;
            continue;
        }
        void *pk = nullptr, *pv = nullptr;
        CU(cudaIpcOpenMemHandle(&pk, kh[r], cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&pv, vh[r], cudaIpcMemLazyEnablePeerAccess));
        c->peer_open[2 * r] = pk;
        c->peer_open[2 * r + 1] = pv;
        c->peers.key[r] = (const unsigned long long*)pk;
        c->peers.vis[r] = (const uint32_t*)pv;
    }
    c->peers.n = n;
    c->peer_rank = my_rank;
    return TRB_OK;
}
int trb_open_peers_raw(TrbCtx* c, const uint64_t* key_ptrs, const uint64_t* vis_ptrs, int n, int my_rank) {
    NOT_WHILE_RECORDING(c, "open_peers_raw");
    if (!c || !c->in_frame || !key_ptrs || !vis_ptrs || n < 1 || n > MAX_PEERS || my_rank < 0 || my_rank >= n)
        return fail(c, TRB_E_COMM, "open_peers_raw: bad argument");
    trb_ipc_close_peers(c);
    for (int r = 0; r < n; ++r) {
        c->peers.key[r] = (const unsigned long long*)(uintptr_t)key_ptrs[r];
        c->peers.vis[r] = (const uint32_t*)(uintptr_t)vis_ptrs[r];
        c->peers.touched[r] = nullptr;
    }
    c->peers.n = n;
    c->peer_rank = my_rank;
    return TRB_OK;
}
int trb_composite_shade_p2p(TrbCtx* c, int y0, int y1) {
    NOT_WHILE_RECORDING(c, "composite_shade_p2p");
    if (!c || !c->in_frame || c->frame.nviews != 1 || c->peers.n < 1 || y0 < 0 || y1 < y0 || y1 > c->frame.H)
        return fail(c, TRB_E_COMM, "composite_shade_p2p: open the peers first (single-view frame)");
    int rc = check_device(c);
    if (rc) return rc;
    rc = enqueue_composite_shade(c, y0, y1);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));   // peers wait on a host barrier after this call
    return TRB_OK;
}

// ---- composite groups: the C face of the sort-last composite (SURVEY 8b trb_comm_init / trb_composite) ------------
int trb_comm_close(TrbCtx* c) {
    if (!c) return TRB_E_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    TrbCtx::Comm& m = c->comm;
    for (void*& p : m.opened)
        if (p) {
            cudaIpcCloseMemHandle(p);
            p = nullptr;
        }
    if (m.my_flags) cudaFree(m.my_flags);
    if (m.ev_drawn) cudaEventDestroy(m.ev_drawn);
    if (m.ev_done) cudaEventDestroy(m.ev_done);
    m = TrbCtx::Comm();
    c->peers.n = 0;
    c->peer_rank = -1;
    c->host_total[8] = 0;
    return TRB_OK;
}
// the rank's "touched" map lives in the same exported 2 MB block as its counters, 4 KB in: one byte per COMPOSITE_CHUNK
// pixels (nullptr when the frame is too large for the block: the composite then reads every rank's keys)
constexpr size_t COMM_BLOCK_BYTES = (size_t)2 << 20, COMM_TOUCHED_OFFSET = 4096;
static const uint8_t* comm_touched_map(const CommFlags* flags, unsigned long long npix) {
    const unsigned long long chunks = (npix + COMPOSITE_CHUNK - 1) / COMPOSITE_CHUNK;
    if (!flags || chunks > COMM_BLOCK_BYTES - COMM_TOUCHED_OFFSET) return nullptr;
    return reinterpret_cast<const uint8_t*>(flags) + COMM_TOUCHED_OFFSET;
}
// mark the chunks of the local key plane this rank drew into (queued behind its draws, in front of "drawn")
static int comm_mark_touched(TrbCtx* c) {
    uint8_t* map = const_cast<uint8_t*>(comm_touched_map(c->comm.my_flags, c->frame.npix));
    if (!map) return TRB_OK;
    const unsigned long long chunks = (c->frame.npix + COMPOSITE_CHUNK - 1) / COMPOSITE_CHUNK;
    Launch L(c, "k_chunk_touched");
    k_chunk_touched<<<(unsigned)((chunks + TPB / 32 - 1) / (TPB / 32)), TPB, 0, c->stream>>>(c->frame.zkey, c->frame.npix, map);
    CU(cudaGetLastError());
    return TRB_OK;
}
static int comm_prepare(TrbCtx* c, int n, int rank) {
    if (!c->in_frame || c->frame.nviews != 1) return fail(c, TRB_E_COMM, "composite group: begin a single-view frame of the final size first");
    int rc = check_device(c);
    if (rc) return rc;
    trb_comm_close(c);
    TrbCtx::Comm& m = c->comm;
    CU(cudaMalloc((void**)&m.my_flags, COMM_BLOCK_BYTES));           // its own block: the IPC handle exports nothing else
    CU(cudaMemsetAsync(m.my_flags, 0, sizeof(CommFlags), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    m.n = n;
    m.rank = rank;
    m.W = c->frame.W;
    m.H = c->frame.H;
    m.key_at_init = c->zkey.p;
    m.peer_flags[rank] = m.my_flags;
    c->peers.key[rank] = c->frame.zkey;
    c->peers.vis[rank] = c->frame.vis;
    c->peers.touched[rank] = comm_touched_map(m.my_flags, c->frame.npix);
    c->peers.n = n;
    c->peer_rank = rank;
    return TRB_OK;
}
int trb_comm_init(TrbCtx* const* ctxs, int n) {
    if (!ctxs || n < 1 || n > MAX_PEERS) return TRB_E_ARG;
    for (int i = 0; i < n; ++i)
        if (!ctxs[i]) return TRB_E_ARG;
    bool one_device_twice = false;
    for (int i = 0; i < n; ++i) {
        TrbCtx* c = ctxs[i];
        int rc = comm_prepare(c, n, i);
        if (rc) return rc;
        if (c->frame.W != ctxs[0]->frame.W || c->frame.H != ctxs[0]->frame.H) return fail(c, TRB_E_COMM, "comm_init: frames differ in size");
        for (int j = 0; j < i; ++j) one_device_twice |= ctxs[j]->device == c->device;
    }
    for (int i = 0; i < n; ++i) {
        TrbCtx* c = ctxs[i];
        CU(cudaSetDevice(c->device));
        c->comm.in_process = true;
        for (int j = 0; j < n; ++j) {
            c->comm.members[j] = ctxs[j];
            if (j == i) continue;
            if (ctxs[j]->device != c->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, c->device, ctxs[j]->device));
                if (!can) return fail(c, TRB_E_COMM, "comm_init: no peer access between the devices of the group");
                cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
                (void)cudaGetLastError();
            }
            c->peers.key[j] = ctxs[j]->frame.zkey;
            c->peers.vis[j] = ctxs[j]->frame.vis;
            c->peers.touched[j] = comm_touched_map(ctxs[j]->comm.my_flags, ctxs[j]->frame.npix);
            c->comm.peer_flags[j] = ctxs[j]->comm.my_flags;
        }
        if (one_device_twice) {     // a waiting kernel could starve the peer it waits for on the same device: events instead
            CU(cudaEventCreateWithFlags(&c->comm.ev_drawn, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->comm.ev_done, cudaEventDisableTiming));
        }
    }
    return TRB_OK;
}
int trb_comm_export(TrbCtx* c, void* blob, size_t blob_bytes) {
    NOT_WHILE_RECORDING(c, "comm_export");
    if (!c || !blob || blob_bytes < TRB_COMM_BLOB_BYTES) return fail(c, TRB_E_ARG, "comm_export: blob too small");
    if (c->comm.n == 0) {                      // first the local half: flags, frame geometry (rank / size follow in comm_open)
        int rc = comm_prepare(c, 1, 0);
        if (rc) return rc;
    }
    int rc = check_device(c);
    if (rc) return rc;
    unsigned char* b = (unsigned char*)blob;
    memset(b, 0, TRB_COMM_BLOB_BYTES);
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)b, c->zkey.p));
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)(b + 64), c->vis.p));
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)(b + 128), c->comm.my_flags));
    int32_t meta[4] = {c->frame.W, c->frame.H, c->device, 0};
    memcpy(b + 192, meta, sizeof(meta));
    return TRB_OK;
}
int trb_comm_open(TrbCtx* c, const void* blobs, int n, int rank) {
    NOT_WHILE_RECORDING(c, "comm_open");
    if (!c || !blobs || n < 1 || n > MAX_PEERS || rank < 0 || rank >= n) return fail(c, TRB_E_ARG, "comm_open: bad argument");
    if (c->comm.n == 0 || !c->comm.my_flags) return fail(c, TRB_E_COMM, "comm_open: call trb_comm_export on this context first");
    int rc = check_device(c);
    if (rc) return rc;
    TrbCtx::Comm& m = c->comm;
    CommFlags* mine = m.my_flags;
    // keep the flags that were exported; forget any earlier peers
    for (void*& p : m.opened)
        if (p) {
            cudaIpcCloseMemHandle(p);
            p = nullptr;
        }
    m.n = n;
    m.rank = rank;
    m.in_process = false;
    m.seq = 0;
    const unsigned char* b = (const unsigned char*)blobs;
    for (int r = 0; r < n; ++r) {
        const unsigned char* br = b + (size_t)r * TRB_COMM_BLOB_BYTES;
        int32_t meta[4];
        memcpy(meta, br + 192, sizeof(meta));
        if (meta[0] != m.W || meta[1] != m.H) return fail(c, TRB_E_COMM, "comm_open: a peer renders a frame of another size");
        if (r == rank) {
            c->peers.key[r] = c->frame.zkey;
            c->peers.vis[r] = c->frame.vis;
            c->peers.touched[r] = comm_touched_map(mine, c->frame.npix);
            m.peer_flags[r] = mine;
            continue;
        }
        void *pk = nullptr, *pv = nullptr, *pf = nullptr;
        CU(cudaIpcOpenMemHandle(&pk, *(const cudaIpcMemHandle_t*)br, cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&pv, *(const cudaIpcMemHandle_t*)(br + 64), cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&pf, *(const cudaIpcMemHandle_t*)(br + 128), cudaIpcMemLazyEnablePeerAccess));
        m.opened[3 * r] = pk; m.opened[3 * r + 1] = pv; m.opened[3 * r + 2] = pf;
        c->peers.key[r] = (const unsigned long long*)pk;
        c->peers.vis[r] = (const uint32_t*)pv;
        c->peers.touched[r] = comm_touched_map((const CommFlags*)pf, c->frame.npix);
        m.peer_flags[r] = (const CommFlags*)pf;
    }
    c->peers.n = n;
    c->peer_rank = rank;
    return TRB_OK;
}
int trb_comm_shard(TrbCtx* c, uint64_t total, uint64_t* first, uint64_t* count) {
    if (!c || c->comm.n < 1 || !first || !count) return fail(c, TRB_E_COMM, "comm_shard: no composite group");
    const uint64_t n = (uint64_t)c->comm.n, r = (uint64_t)c->comm.rank, base = total / n, rem = total % n;
    *first = r * base + std::min(r, rem);
    *count = base + (r < rem ? 1 : 0);
    return TRB_OK;
}
int trb_comm_rows(TrbCtx* c, int* y0, int* y1) {
    if (!c || c->comm.n < 1 || !y0 || !y1) return fail(c, TRB_E_COMM, "comm_rows: no composite group");
    comm_rows(c->comm.H, c->comm.rank, c->comm.n, y0, y1);
    return TRB_OK;
}
// One rank's half of a frame's composite, entirely on its stream: publish "my draws are complete", wait for the peers'
// draws, composite + shade the rows this rank owns, publish "I have finished reading".  Nothing here waits on the host.
int trb_composite(TrbCtx* c) {
    NOT_WHILE_RECORDING(c, "composite");
    if (!c || !c->in_frame || c->frame.nviews != 1 || c->comm.n < 1) return fail(c, TRB_E_COMM, "composite: no composite group (trb_comm_init / trb_comm_open)");
    if (c->comm.in_process && c->comm.ev_drawn)
        return fail(c, TRB_E_COMM, "composite: contexts that share a device are composited together, with trb_composite_group");
    int rc = check_device(c);
    if (rc) return rc;
    TrbCtx::Comm& m = c->comm;
    ++m.seq;
    rc = comm_mark_touched(c);
    if (rc) return rc;
    {
        Launch L(c, "k_comm_publish");
        k_comm_publish<<<1, 1, 0, c->stream>>>(&m.my_flags->drawn, m.seq);
    }
    rc = comm_wait_peers(c, /*done=*/false);
    if (rc) return rc;
    int y0, y1;
    comm_rows(m.H, m.rank, m.n, &y0, &y1);
    rc = enqueue_composite_shade(c, y0, y1);
    if (rc) return rc;
    {
        Launch L(c, "k_comm_publish");
        k_comm_publish<<<1, 1, 0, c->stream>>>(&m.my_flags->done, m.seq);
    }
    CU(cudaGetLastError());
    return TRB_OK;
}
// The same for every context of an in-process group at once (one host thread drives all of them; required when
// contexts share a device): events order the streams.
int trb_composite_group(TrbCtx* const* ctxs, int n) {
    if (!ctxs || n < 1 || n > MAX_PEERS) return TRB_E_ARG;
    for (int i = 0; i < n; ++i) {
        TrbCtx* c = ctxs[i];
        if (!c || !c->in_frame || c->frame.nviews != 1 || c->comm.n != n || c->comm.rank != i || !c->comm.in_process)
            return fail(c, TRB_E_COMM, "composite_group: pass the contexts of trb_comm_init in the same order");
    }
    for (int i = 0; i < n; ++i) {          // everybody's draws are queued: mark them
        TrbCtx* c = ctxs[i];
        CU(cudaSetDevice(c->device));
        ++c->comm.seq;
        int rc = comm_mark_touched(c);
        if (rc) return rc;
        if (c->comm.ev_drawn) CU(cudaEventRecord(c->comm.ev_drawn, c->stream));
        else {
            Launch L(c, "k_comm_publish");
            k_comm_publish<<<1, 1, 0, c->stream>>>(&c->comm.my_flags->drawn, c->comm.seq);
        }
    }
    for (int i = 0; i < n; ++i) {
        TrbCtx* c = ctxs[i];
        CU(cudaSetDevice(c->device));
        int rc = comm_wait_peers(c, /*done=*/false);
        if (rc) return rc;
        int y0, y1;
        comm_rows(c->comm.H, i, n, &y0, &y1);
        rc = enqueue_composite_shade(c, y0, y1);
        if (rc) return rc;
        if (c->comm.ev_done) CU(cudaEventRecord(c->comm.ev_done, c->stream));
        else {
            Launch L(c, "k_comm_publish");
            k_comm_publish<<<1, 1, 0, c->stream>>>(&c->comm.my_flags->done, c->comm.seq);
        }
        CU(cudaGetLastError());
    }
    return TRB_OK;
}
int trb_set_shade_rows(TrbCtx* c, int y0, int y1) {
    NOT_WHILE_RECORDING(c, "set_shade_rows");
    if (!c || !c->in_frame || y0 < 0 || y1 < y0 || y1 > c->frame.H) return fail(c, TRB_E_ARG, "set_shade_rows");
    c->shade_row0 = y0;
    c->shade_row1 = y1;
    return TRB_OK;
}

// ---- host helpers (reference operation order, no device work) -------------------------------------
void trb_light_dir_eye(const double* mv, const double* dir, double* out) {
    // main.cpp:59-68: mat<3,3> * vec3 (dot<3> from +0.0) then normalized()
    D3 d{dir[0], dir[1], dir[2]};
    D3 r{dot3(D3{mv[0], mv[1], mv[2]}, d), dot3(D3{mv[4], mv[5], mv[6]}, d), dot3(D3{mv[8], mv[9], mv[10]}, d)};
    r = normalize3(r);
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
}
void trb_mat4_mul_batch(const double* a, int n, const double* b, double* out) {
    for (int v = 0; v < n; ++v) trb_mat4_mul(a + 16 * v, b, out + 16 * v);
}
void trb_light_dir_eye_batch(const double* mvs, int n, const double* dir, double* out) {
    for (int v = 0; v < n; ++v) trb_light_dir_eye(mvs + 16 * v, dir, out + 3 * v);
}
void trb_lookat(const double* eye, const double* center, const double* up, double* out) {
    // our_gl.cpp:25-41
    D3 e{eye[0], eye[1], eye[2]}, ce{center[0], center[1], center[2]}, u{up[0], up[1], up[2]};
    D3 z = normalize3(sub3(e, ce));
    D3 x = normalize3(D3{u.y * z.z - u.z * z.y, u.z * z.x - u.x * z.z, u.x * z.y - u.y * z.x});
    D3 y{z.y * x.z - z.z * x.y, z.z * x.x - z.x * x.z, z.x * x.y - z.y * x.x};
    for (int i = 0; i < 16; ++i) out[i] = (i % 5 == 0) ? 1.0 : 0.0;
    out[0] = x.x; out[1] = x.y; out[2] = x.z; out[3] = -dot3(x, e);
    out[4] = y.x; out[5] = y.y; out[6] = y.z; out[7] = -dot3(y, e);
    out[8] = z.x; out[9] = z.y; out[10] = z.z; out[11] = -dot3(z, e);
}
void trb_perspective(double fov_deg, double aspect, double zn, double zf, double* out) {
    // our_gl.cpp:44-56
    double fov_rad = fov_deg * M_PI / 180.0;
    double t = tan(fov_rad / 2.0);
    for (int i = 0; i < 16; ++i) out[i] = (i % 5 == 0) ? 1.0 : 0.0;
    out[0] = 1.0 / (aspect * t);
    out[5] = 1.0 / t;
    out[10] = (zf + zn) / (zn - zf);
    out[11] = (2.0 * zf * zn) / (zn - zf);
    out[14] = -1.0;
    out[15] = 0.0;
}
void trb_viewport(int x, int y, int w, int h, double* out) {
    // our_gl.cpp:59-69
    for (int i = 0; i < 16; ++i) out[i] = (i % 5 == 0) ? 1.0 : 0.0;
    out[0] = w / 2.0;
    out[5] = h / 2.0;
    out[3] = x + w / 2.0;
    out[7] = y + h / 2.0;
    out[10] = 1.0;
    out[11] = 0.0;
}
void trb_frustum_planes(const double* m, double* planes) {
    // our_gl.cpp:212-261: plane k reads m[r][3] +- m[r][col] for r = 0..2 and m[3][3] +- m[3][col] (sic: transposed)
    static const int col[6] = {0, 0, 1, 1, 2, 2};
    static const double sgn[6] = {1, -1, 1, -1, 1, -1};
    for (int p = 0; p < 6; ++p) {
        D3 n{m[3] + sgn[p] * m[col[p]], m[7] + sgn[p] * m[4 + col[p]], m[11] + sgn[p] * m[8 + col[p]]};
        double d = m[15] + sgn[p] * m[12 + col[p]];
        const double len = sqrt(dot3(n, n));                     // norm(), geometry.h:131-133
        if (len > 0.0) {
            n = D3{n.x / len, n.y / len, n.z / len};             // vec / double, geometry.h:113-118
            d /= len;
        }
        planes[4 * p] = n.x; planes[4 * p + 1] = n.y; planes[4 * p + 2] = n.z; planes[4 * p + 3] = d;
    }
}
int trb_frustum_intersects(const double* planes, const double* lo, const double* hi) {
    // our_gl.cpp:263-280
    for (int p = 0; p < 6; ++p) {
        const double* pl = planes + 4 * p;
        D3 far_corner{lo[0], lo[1], lo[2]};
        if (pl[0] >= 0) far_corner.x = hi[0];
        if (pl[1] >= 0) far_corner.y = hi[1];
        if (pl[2] >= 0) far_corner.z = hi[2];
        if (dot3(D3{pl[0], pl[1], pl[2]}, far_corner) + pl[3] < 0) return 0;   // Plane::distance, geometry.h:266-268
    }
    return 1;
}
void trb_aabb_transform(const double* lo, const double* hi, const double* m, double* out_lo, double* out_hi) {
    // geometry.h:297-327
    double nlo[3] = {1e9, 1e9, 1e9}, nhi[3] = {-1e9, -1e9, -1e9};
    for (int i = 0; i < 8; ++i) {
        const double cx = (i & 1) ? hi[0] : lo[0], cy = (i & 2) ? hi[1] : lo[1], cz = (i & 4) ? hi[2] : lo[2];
        const double t[4] = {dot4(m, cx, cy, cz, 1.0), dot4(m + 4, cx, cy, cz, 1.0), dot4(m + 8, cx, cy, cz, 1.0),
                             dot4(m + 12, cx, cy, cz, 1.0)};
        for (int k = 0; k < 3; ++k) {
            const double v = t[k] / t[3];
            nlo[k] = std::min(nlo[k], v);
            nhi[k] = std::max(nhi[k], v);
        }
    }
    for (int k = 0; k < 3; ++k) { out_lo[k] = nlo[k]; out_hi[k] = nhi[k]; }
}
void trb_cull_batch(const double* perspective, const double* views, int n, const double* lo, const double* hi, uint8_t* out) {
    for (int v = 0; v < n; ++v) {
        double vp[16], planes[24];
        trb_mat4_mul(perspective, views + 16 * v, vp);           // main.cpp:623
        trb_frustum_planes(vp, planes);
        out[v] = (uint8_t)trb_frustum_intersects(planes, lo, hi);
    }
}
void trb_mat4_mul(const double* a, const double* b, double* out) {
    // geometry.h:195-205
    double r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            r[i * 4 + j] = s;
        }
    memcpy(out, r, sizeof(r));
}

}  // extern "C"
