// Peer-memory data movement for the column-sharded LR-ADI: the kernels below read
// and write buffers of OTHER GPUs through pointers mapped into this process (symmetric memory,
// CUDA VMM over NVLink / NVSwitch), so that NCCL is not needed for the bulk traffic:
//   put2d      2-D block copy into a peer's buffer (P2P stores): the column -> row re-shard of
//              the factor before the Gram product, and the all-gather of the compressed rows
//   sum_peers  out[e] = sum over ranks g (fixed order) of peer_g[e] (P2P loads): the K x K Gram
//              all-reduce as the epilogue of the local partial products - every rank reads the
//              same numbers in the same order, so the result is bitwise identical everywhere.
// Ordering between the ranks (a peer must have finished writing before we read) is the
// caller's: a symmetric-memory barrier between the phases.
#include "common.cuh"
#include <algorithm>

namespace ocb {

__global__ void __launch_bounds__(256) put2d_kernel(const double* __restrict__ src, int64_t lds,
                                                   int64_t nrows, int64_t ncols,
                                                   double* __restrict__ dst, int64_t ldd) {
    const int64_t total = nrows * ncols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / ncols, c = e - r * ncols;
        dst[r * ldd + c] = src[r * lds + c];
    }
}

constexpr int P2P_MAX_RANKS = 16;
struct PeerPtrs {
    const double* p[P2P_MAX_RANKS];
};

__global__ void __launch_bounds__(256) sum_peers_kernel(const PeerPtrs peers, int world, int64_t count,
                                                       double* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count;
         e += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int g = 0; g < world; ++g) s += peers.p[g][e];
        out[e] = s;
    }
}

}  // namespace ocb

extern "C" {

int ocb_p2p_put2d(const double* d_src, int64_t lds, int64_t nrows, int64_t ncols, double* d_dst_peer,
                  int64_t ldd, void* stream) {
    using namespace ocb;
    OCB_ARG(nrows >= 0 && ncols >= 0 && lds >= ncols && ldd >= ncols, "p2p_put2d sizes");
    if (nrows == 0 || ncols == 0) return OCB_OK;
    OCB_ARG(d_src && d_dst_peer, "p2p_put2d null");
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((nrows * ncols + 1023) / 1024, 148 * 8));
    put2d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_src, lds, nrows, ncols, d_dst_peer, ldd);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

int ocb_p2p_sum_peers(const double* const* h_peer_ptrs, int64_t world, int64_t count, double* d_out,
                      void* stream) {
    using namespace ocb;
    OCB_ARG(h_peer_ptrs && world >= 1 && world <= P2P_MAX_RANKS && count >= 0 && d_out, "p2p_sum_peers");
    if (count == 0) return OCB_OK;
    PeerPtrs pp;
    for (int g = 0; g < P2P_MAX_RANKS; ++g) pp.p[g] = g < world ? h_peer_ptrs[g] : nullptr;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((count + 1023) / 1024, 148 * 8));
    sum_peers_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pp, (int)world, count, d_out);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}
}
