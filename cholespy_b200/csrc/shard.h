// RHS-column sharding of one solver over several B200s INSIDE the solve call (single process).
//
// New: the reference is single-device (it always uses device 0, reference src/cuda_driver.cpp:124-125); BASELINE.json's
// north_star asks the GPU branch to "shard multi-RHS work across GPUs".  The path shards only across right-hand-side
// columns (rows are coupled by the triangular dependency), so:
//   * once: the device image (both sweep slabs + task tables) of the primary solver is replicated to every other
//     device -- ncclBroadcast over NVLink (libnccl.so.2 is dlopen'ed; one communicator per device, ncclCommInitAll,
//     grouped calls), or peer copies (cuMemcpyPeerAsync) where NCCL is not available;
//   * per solve: device d gets a contiguous block of columns.  The primary solves its block IN PLACE in the caller's
//     b / x (row stride = the full width); every other device receives its block as a packed (n, k_d) array by a
//     strided peer copy (cuMemcpy2DAsync over NVLink), runs the two sweeps on its own legacy stream, and its x block
//     is copied back the same way.  Event joins make the whole thing one asynchronous operation on the PRIMARY's
//     legacy stream: the copies wait for an event recorded there at entry, and that stream waits for every device's
//     completion event before the call returns to the caller -- the semantics of the unsharded call
//     ("make them wait on an event recorded on the legacy stream ... join before returning").
// No data-path collective: columns are independent.
#pragma once
#include <string>
#include <vector>

#include "gpu_solver.h"

namespace chb {

class ShardGroup {
public:
    // `primary` stays owned by the caller and must outlive the group.  `devices`: ordinals of the ADDITIONAL devices.
    ShardGroup(GpuSolver *primary, const std::vector<int> &devices);
    ~ShardGroup();
    ShardGroup(const ShardGroup &) = delete;
    ShardGroup &operator=(const ShardGroup &) = delete;

    int n_devices() const { return 1 + (int) m_remote.size(); }
    const char *replication() const { return m_replication.c_str(); }  // "nccl" or "peer"
    double replication_seconds() const { return m_replication_s; }
    size_t replicated_bytes() const { return m_replicated_bytes; }
    // b, x: (n, k) row-major arrays on the primary device.  Fewer than min_columns_per_device columns per device:
    // fewer devices are used (k = 1..3 stays on the primary).
    void solve(int k, const void *b, void *x);
    // [(first column, columns)] of the split solve() uses for k columns
    std::vector<std::pair<int, int>> split(int k) const;

private:
    struct Remote {
        DeviceState *dev = nullptr;
        GpuSolver *solver = nullptr;
        CUstream stream = nullptr;        // copy stream in the remote context
        CUevent ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
        void *stage_b = nullptr, *stage_x = nullptr;
        size_t stage_bytes = 0;
    };
    void replicate_nccl(const std::vector<int> &ordinals);
    void replicate_peer();

    GpuSolver *m_primary;
    DeviceState *m_dev0;
    std::vector<Remote> m_remote;
    CUevent m_ev_begin = nullptr;
    std::string m_replication;
    double m_replication_s = 0.0;
    size_t m_replicated_bytes = 0;
    int m_min_cols = 8;
};

}  // namespace chb
