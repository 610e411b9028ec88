// peer_kernels.cuh — global-qubit remap of a sharded state over NVLink peer memory.
//
// A state of n qubits is sharded over R = 2^g GPUs: the top g index bits are the rank
// number.  A non-diagonal gate on a "global" qubit first swaps that qubit with a local
// one.  Swapping k (global bit, local bit) pairs at once is an INVOLUTION on the
// amplitudes: the element held by rank-group member r at a local index whose selected
// local bits read v trades places with the element of member v whose selected bits read
// r (and stays put when v == r).  So the whole remap runs IN PLACE with no staging
// buffer: every unordered pair of elements is owned by exactly one thread of one GPU,
// which loads both (one from its own HBM, one from the peer's over NVLink) and stores
// both.  The two members of a rank pair split their common pairs in half by one more
// local bit (`hbit`), so that each NVLink direction carries the same number of read
// responses and writes.  NVLink bytes per GPU and direction: (2^k - 1) / 2^(k+1) of a
// shard, instead of k/2 of a shard (plus a staging copy each) for k sequential
// half-shard send/recv exchanges.
//
// The caller orders the kernel between two cross-rank barriers on the stream (every
// peer has finished its previous kernels; nobody reads the shard before all swaps are
// done) — afquantumsim_b200/sharded.py uses stream-ordered NCCL all_reduce calls.
#pragma once
#include "common.cuh"

namespace aqs {

constexpr int kPeerMaxK = 3;                 // up to 8 GPUs in one remap
constexpr int kPeerThreads = 256;
constexpr int kPeerItems = 4;                // independent 128-bit pairs in flight per thread

struct PeerSwapArgs {
    float2* mine;
    float2* peer[1 << kPeerMaxK];            // shard of group member v (peer[my] unused)
    uint64_t n_items;                        // 128-bit work items per partner
    uint64_t voff[1 << kPeerMaxK];           // value v deposited on the selected local bits
    uint32_t my;                             // this GPU's value on the selected global bits
    uint32_t k;
    uint32_t hbit;                           // local bit that splits a rank pair's work
    BitList fixed;                           // bit 0 (vector), the selected local bits, hbit
};

__global__ void __launch_bounds__(kPeerThreads) k_peer_bitswap(const __grid_constant__ PeerSwapArgs P) {
    // Partner of this CTA: member my ^ d, d = 1 .. 2^k - 1, with d the FASTEST-varying part of the block number, so that at
    // any moment every GPU talks to all its partners at once and, for each d, the pairs (r, r ^ d) form a perfect matching:
    // no GPU's NVLink ingress becomes the hot spot.  (Enumerating the partners one after the other — every rank starting
    // with member 0 — made the 8-GPU remap of brickwork-34 take 51 ms for 7 GiB each way: 38 % of the link.)
    const uint32_t n_partners = (1u << P.k) - 1u;
    const uint32_t v = P.my ^ (blockIdx.x % n_partners + 1u);
    const uint64_t bx = blockIdx.x / n_partners;
    float2* __restrict__ other = P.peer[v];
    const uint64_t hsel = (P.my < v) ? 0ull : (1ull << P.hbit);
    const uint64_t off_mine = P.voff[v] | hsel;        // my element: selected bits read v
    const uint64_t off_peer = P.voff[P.my] | hsel;     // the partner's element: selected bits read my value
    const uint64_t j0 = bx * (kPeerThreads * kPeerItems) + threadIdx.x;
    float4 a[kPeerItems], b[kPeerItems];
    uint64_t base[kPeerItems];
#pragma unroll
    for (int u = 0; u < kPeerItems; ++u) {
        const uint64_t j = j0 + (uint64_t)u * kPeerThreads;
        base[u] = deposit_zeros(j, P.fixed);
        if (j < P.n_items) {
            b[u] = *reinterpret_cast<const float4*>(other + (base[u] | off_peer));   // remote read first: longest latency
            a[u] = *reinterpret_cast<const float4*>(P.mine + (base[u] | off_mine));
        }
    }
#pragma unroll
    for (int u = 0; u < kPeerItems; ++u) {
        const uint64_t j = j0 + (uint64_t)u * kPeerThreads;
        if (j < P.n_items) {
            *reinterpret_cast<float4*>(P.mine + (base[u] | off_mine)) = b[u];
            *reinterpret_cast<float4*>(other + (base[u] | off_peer)) = a[u];
        }
    }
}

}  // namespace aqs
