// RHS-column sharding over several devices inside one solve call.  See shard.h.
#include "shard.h"

#include <dlfcn.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace chb {

namespace {

struct ContextGuard {
    explicit ContextGuard(DeviceState *d) { cuda_check(cuCtxPushCurrent(d->ctx)); }
    ~ContextGuard() { cuda_check(cuCtxPopCurrent(nullptr)); }
};

void check(CUresult rc, const char *what) {
    if (rc != CUDA_SUCCESS) {
        cuda_check(rc);
        throw std::runtime_error(std::string("multi-GPU sharding: ") + what + " failed: " + driver_error_text());
    }
}

// -- NCCL through dlopen (no link-time dependency; absent library => peer copies) -------------------------------------
using ncclComm_t = struct ncclComm *;
struct Nccl {
    void *lib = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, CUstream) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
Nccl &nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char *disable = std::getenv("CHB_SHARD_NO_NCCL");
    if (disable && std::atoi(disable) != 0) return n;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        n.lib = dlopen(name, RTLD_LAZY | RTLD_GLOBAL);
        if (n.lib) break;
    }
    if (!n.lib) return n;
    n.CommInitAll = (decltype(n.CommInitAll)) dlsym(n.lib, "ncclCommInitAll");
    n.CommDestroy = (decltype(n.CommDestroy)) dlsym(n.lib, "ncclCommDestroy");
    n.GroupStart = (decltype(n.GroupStart)) dlsym(n.lib, "ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd)) dlsym(n.lib, "ncclGroupEnd");
    n.Broadcast = (decltype(n.Broadcast)) dlsym(n.lib, "ncclBroadcast");
    n.GetErrorString = (decltype(n.GetErrorString)) dlsym(n.lib, "ncclGetErrorString");
    n.ok = n.CommInitAll && n.CommDestroy && n.GroupStart && n.GroupEnd && n.Broadcast;
    return n;
}
constexpr int kNcclUint8 = 1;  // ncclUint8 (nccl.h)

}  // namespace

ShardGroup::ShardGroup(GpuSolver *primary, const std::vector<int> &devices) : m_primary(primary) {
    m_dev0 = device_state(primary->device());
    if (!m_dev0) throw std::runtime_error(driver_error_text());
    if (const char *v = std::getenv("CHB_SHARD_MIN_COLUMNS")) m_min_cols = std::max(1, std::atoi(v));
    const ImageShape shape = primary->shape();
    std::vector<int> ordinals{primary->device()};
    for (int d : devices) {
        if (d == primary->device()) continue;
        DeviceState *ds = device_state(d);
        if (!ds) throw std::runtime_error(std::string("multi-GPU sharding: ") + driver_error_text());
        // peer access both ways (the strided column-block copies go straight over NVLink)
        int can = 0;
        check(cuDeviceCanAccessPeer(&can, m_dev0->dev, ds->dev), "cuDeviceCanAccessPeer");
        if (can) {
            {
                ContextGuard g(m_dev0);
                const CUresult rc = cuCtxEnablePeerAccess(ds->ctx, 0);
                if (rc != CUDA_SUCCESS && rc != 704 /* already enabled */) check(rc, "cuCtxEnablePeerAccess");
            }
            {
                ContextGuard g(ds);
                const CUresult rc = cuCtxEnablePeerAccess(m_dev0->ctx, 0);
                if (rc != CUDA_SUCCESS && rc != 704) check(rc, "cuCtxEnablePeerAccess");
            }
        }
        Remote r;
        r.dev = ds;
        r.solver = new GpuSolver(d, primary->elem_size(), shape);
        {
            ContextGuard g(ds);
            check(cuStreamCreate(&r.stream, 1), "cuStreamCreate");  // CU_STREAM_NON_BLOCKING
            check(cuEventCreate(&r.ev_in, 2), "cuEventCreate");     // CU_EVENT_DISABLE_TIMING
            check(cuEventCreate(&r.ev_done, 2), "cuEventCreate");
            check(cuEventCreate(&r.ev_out, 2), "cuEventCreate");
        }
        m_remote.push_back(r);
        ordinals.push_back(d);
    }
    {
        ContextGuard g(m_dev0);
        check(cuEventCreate(&m_ev_begin, 2), "cuEventCreate");
    }
    const auto t0 = std::chrono::steady_clock::now();
    if (!m_remote.empty()) {
        if (nccl().ok) {
            try {
                replicate_nccl(ordinals);
                m_replication = "nccl";
            } catch (const std::exception &) {
                m_replication.clear();
            }
        }
        if (m_replication.empty()) {
            replicate_peer();
            m_replication = "peer";
        }
    }
    m_replication_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

ShardGroup::~ShardGroup() {
    for (Remote &r : m_remote) {
        {
            ContextGuard g(r.dev);
            cuda_check(cuStreamSynchronize(r.stream));
            cuda_check(cuStreamSynchronize(nullptr));
            if (r.stage_b) cuda_check(cuMemFree(r.stage_b));
            if (r.stage_x) cuda_check(cuMemFree(r.stage_x));
            cuda_check(cuEventDestroy(r.ev_in));
            cuda_check(cuEventDestroy(r.ev_done));
            cuda_check(cuEventDestroy(r.ev_out));
            cuda_check(cuStreamDestroy(r.stream));
        }
        delete r.solver;
    }
    if (m_ev_begin) {
        ContextGuard g(m_dev0);
        cuda_check(cuEventDestroy(m_ev_begin));
    }
}

// One communicator per device (ncclCommInitAll); every image buffer is one grouped ncclBroadcast from the primary.
void ShardGroup::replicate_nccl(const std::vector<int> &ordinals) {
    Nccl &n = nccl();
    const int nd = (int) ordinals.size();
    std::vector<ncclComm_t> comms((size_t) nd, nullptr);
    int rc = n.CommInitAll(comms.data(), nd, ordinals.data());
    if (rc != 0) throw std::runtime_error(std::string("ncclCommInitAll: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error"));
    const std::vector<ImageBuffer> src = m_primary->image();
    bool failed = false;
    for (size_t bi = 0; bi < src.size() && !failed; ++bi) {
        rc = n.GroupStart();
        for (int d = 0; d < nd && rc == 0; ++d) {
            DeviceState *ds = d == 0 ? m_dev0 : m_remote[(size_t) d - 1].dev;
            ContextGuard g(ds);
            void *recv = d == 0 ? src[bi].dptr : m_remote[(size_t) d - 1].solver->image()[bi].dptr;
            rc = n.Broadcast(src[bi].dptr, recv, src[bi].bytes, kNcclUint8, 0, comms[(size_t) d], nullptr);
        }
        const int rc_end = n.GroupEnd();
        failed = rc != 0 || rc_end != 0;
        m_replicated_bytes += src[bi].bytes;
    }
    for (int d = 0; d < nd; ++d) {
        DeviceState *ds = d == 0 ? m_dev0 : m_remote[(size_t) d - 1].dev;
        ContextGuard g(ds);
        cuda_check(cuStreamSynchronize(nullptr));
    }
    for (ncclComm_t c : comms)
        if (c) n.CommDestroy(c);
    if (failed) {
        m_replicated_bytes = 0;
        throw std::runtime_error("ncclBroadcast failed");
    }
}

void ShardGroup::replicate_peer() {
    const std::vector<ImageBuffer> src = m_primary->image();
    m_replicated_bytes = 0;
    for (Remote &r : m_remote) {
        const std::vector<ImageBuffer> dst = r.solver->image();
        ContextGuard g(r.dev);
        for (size_t bi = 0; bi < src.size(); ++bi)
            check(cuMemcpyPeerAsync(dst[bi].dptr, r.dev->ctx, src[bi].dptr, m_dev0->ctx, src[bi].bytes, r.stream), "cuMemcpyPeerAsync");
        check(cuStreamSynchronize(r.stream), "cuStreamSynchronize");
    }
    for (const ImageBuffer &b : src) m_replicated_bytes += b.bytes;
}

std::vector<std::pair<int, int>> ShardGroup::split(int k) const {
    // whole groups of 4 columns (16-byte column offsets for single precision) dealt as evenly as possible
    int nd = std::min(n_devices(), std::max(1, k / m_min_cols));
    std::vector<std::pair<int, int>> out;
    const int groups = (k + 3) / 4;
    nd = std::min(nd, groups);
    int c0 = 0;
    for (int d = 0; d < nd; ++d) {
        const int g = groups / nd + (d < groups % nd ? 1 : 0);
        const int kc = std::min(4 * g, k - c0);
        out.push_back({c0, kc});
        c0 += kc;
    }
    return out;
}

void ShardGroup::solve(int k, const void *b, void *x) {
    if (k > 128 && !GpuSolver::wide_rhs_allowed()) throw std::invalid_argument("The number of RHS should be less than 128.");
    if (k < 1) throw std::invalid_argument("The number of RHS should be at least 1.");
    const std::vector<std::pair<int, int>> blocks = split(k);
    if (blocks.size() <= 1) {
        m_primary->solve_wide(k, b, x);
        return;
    }
    const size_t es = (size_t) m_primary->elem_size();
    const size_t n = (size_t) m_primary->n();
    {
        ContextGuard g(m_dev0);
        check(cuEventRecord(m_ev_begin, nullptr), "cuEventRecord");  // b is final on the primary's legacy stream from here on
    }
    auto copy_block = [&](bool to_remote, Remote &r, int c0, int kc) {
        CUDA_MEMCPY2D d;
        std::memset(&d, 0, sizeof d);
        d.WidthInBytes = es * (size_t) kc;
        d.Height = n;
        d.srcMemoryType = 2, d.dstMemoryType = 2;
        if (to_remote) {
            d.srcDevice = (CUdeviceptr) b, d.srcPitch = es * (size_t) k, d.srcXInBytes = es * (size_t) c0;
            d.dstDevice = r.stage_b, d.dstPitch = es * (size_t) kc;
        } else {
            d.srcDevice = r.stage_x, d.srcPitch = es * (size_t) kc;
            d.dstDevice = x, d.dstPitch = es * (size_t) k, d.dstXInBytes = es * (size_t) c0;
        }
        check(cuMemcpy2DAsync(&d, r.stream), "cuMemcpy2DAsync");
    };
    for (size_t bi = 1; bi < blocks.size(); ++bi) {
        Remote &r = m_remote[bi - 1];
        const int c0 = blocks[bi].first, kc = blocks[bi].second;
        const size_t bytes = es * n * (size_t) kc;
        {
            ContextGuard g(r.dev);
            if (bytes > r.stage_bytes) {
                check(cuStreamSynchronize(r.stream), "cuStreamSynchronize");
                check(cuStreamSynchronize(nullptr), "cuStreamSynchronize");
                if (r.stage_b) cuda_check(cuMemFree(r.stage_b));
                if (r.stage_x) cuda_check(cuMemFree(r.stage_x));
                r.stage_b = r.stage_x = nullptr;
                r.stage_bytes = 0;
                check(cuMemAlloc(&r.stage_b, bytes), "cuMemAlloc");
                check(cuMemAlloc(&r.stage_x, bytes), "cuMemAlloc");
                r.stage_bytes = bytes;
            }
            check(cuStreamWaitEvent(r.stream, m_ev_begin, 0), "cuStreamWaitEvent");
            check(cuStreamWaitEvent(r.stream, r.ev_done, 0), "cuStreamWaitEvent");  // the previous solve's sweeps no longer read stage_b
            copy_block(true, r, c0, kc);
            check(cuEventRecord(r.ev_in, r.stream), "cuEventRecord");
            check(cuStreamWaitEvent(nullptr, r.ev_in, 0), "cuStreamWaitEvent");     // remote legacy stream
            check(cuStreamWaitEvent(nullptr, r.ev_out, 0), "cuStreamWaitEvent");    // the previous copy-back no longer reads stage_x
        }
        r.solver->solve_wide(kc, r.stage_b, r.stage_x);  // enqueued on the remote device's legacy stream
        {
            ContextGuard g(r.dev);
            check(cuEventRecord(r.ev_done, nullptr), "cuEventRecord");
            check(cuStreamWaitEvent(r.stream, r.ev_done, 0), "cuStreamWaitEvent");
            copy_block(false, r, c0, kc);
            check(cuEventRecord(r.ev_out, r.stream), "cuEventRecord");
        }
    }
    // the primary's own block, in place (row stride k)
    m_primary->solve_wide(blocks[0].second, (const char *) b + es * (size_t) blocks[0].first, (char *) x + es * (size_t) blocks[0].first, k, k);
    {
        ContextGuard g(m_dev0);
        for (size_t bi = 1; bi < blocks.size(); ++bi) check(cuStreamWaitEvent(nullptr, m_remote[bi - 1].ev_out, 0), "cuStreamWaitEvent");
    }
}

}  // namespace chb
