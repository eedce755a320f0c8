// In-kernel all-gather over peer-mapped mailboxes (lhn_exchange, include/lhn.h): device-side protocol, shared by the
// persistent heatmap kernel (lhn_heatmap_team.cuh) and the flush kernel (lhn_exchange.cu).
//
// Mailbox of one rank (LHN_XCH_MAILBOX_BYTES, zeroed once, mapped into every peer):
//   [slot 0..3][source rank 0..7] payload of LHN_XCH_PAYLOAD_BYTES      slot = step number & 3
//   control page: flags u32 [slot][source rank] (the step number once the payload has landed),
//                 +1024 launch tickets u32 [slot], +2048 timing stamps u64 [slot][8]
// A step's block travels as ONE bulk copy (TMA, shared -> peer global) per peer, all peers in flight together, then a
// system-scope release of one flag per peer; the receiver polls its own (local) flags with acquire loads.
#pragma once
#include "lhn_common.cuh"

namespace lhn {

struct XchCtx {
  unsigned char* mail[LHN_XCH_MAX_RANKS];   // mailbox of rank r as mapped into this process
  int world, rank;                          // world == 0: no exchange
  unsigned int timeout_ms;
  int* status;
};

__device__ __forceinline__ unsigned char* xch_ctrl(const XchCtx& x, int mailbox_rank) {
  return x.mail[mailbox_rank] + (size_t)LHN_XCH_SLOTS * LHN_XCH_MAX_RANKS * LHN_XCH_PAYLOAD_BYTES;
}
__device__ __forceinline__ unsigned char* xch_slot(const XchCtx& x, unsigned seq, int mailbox_rank, int src_rank) {
  return x.mail[mailbox_rank] + ((size_t)(seq & (LHN_XCH_SLOTS - 1)) * LHN_XCH_MAX_RANKS + src_rank) * LHN_XCH_PAYLOAD_BYTES;
}
__device__ __forceinline__ unsigned int* xch_flag(const XchCtx& x, unsigned seq, int mailbox_rank, int src_rank) {
  return reinterpret_cast<unsigned int*>(xch_ctrl(x, mailbox_rank)) + (seq & (LHN_XCH_SLOTS - 1)) * LHN_XCH_MAX_RANKS + src_rank;
}
// launch ticket: one per slot, so consecutive (overlapping) launches never share one
__device__ __forceinline__ unsigned int* xch_ticket(const XchCtx& x, unsigned seq) {
  return reinterpret_cast<unsigned int*>(xch_ctrl(x, x.rank) + 1024) + (seq & (LHN_XCH_SLOTS - 1));
}
// timing stamps (globaltimer ns) of the exchange of step `seq`: [seq, t_enter, t_published, t_peers_arrived]
__device__ __forceinline__ unsigned long long* xch_stamps(const XchCtx& x, unsigned seq) {
  return reinterpret_cast<unsigned long long*>(xch_ctrl(x, x.rank) + 2048) + 8 * (seq & (LHN_XCH_SLOTS - 1));
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One warp.  `local`: n 64-bit words in global memory (n * 8 <= LHN_XCH_PAYLOAD_BYTES).  `stage`: 16-byte aligned
// shared memory of n + 1 words, or nullptr for blocks of <= 32 words (plain stores).  Sends the block of step `seq`
// to every peer's mailbox and releases the flags; does not wait for anybody.
static __device__ __noinline__ void xch_publish(const XchCtx& x, unsigned seq, const unsigned long long* local, int n,
                                                int lane, unsigned long long* stage) {
  const int world = x.world, me = x.rank;
  unsigned long long* stamps = xch_stamps(x, seq);
  if (lane == 0) { stamps[0] = seq; stamps[1] = gtimer(); }
  if (stage) {
    const int n2 = (n + 1) & ~1;                             // bulk copies move multiples of 16 bytes
    // sixteen independent loads per lane in flight: under the streaming kernels' traffic one L2 round trip costs
    // 1-3 us, and a block of 400 words read four words at a time took 4-12 us (profiles/r02_xch_timing.txt)
    for (int e0 = lane; e0 < n2; e0 += 32 * 16) {
      unsigned long long v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) { const int e = e0 + 32 * u; v[u] = e < n ? __ldcg(local + e) : 0ull; }
#pragma unroll
      for (int u = 0; u < 16; ++u) { const int e = e0 + 32 * u; if (e < n2) stage[e] = v[u]; }
    }
    __syncwarp();
    fence_proxy_async();                                     // generic-proxy writes of the stage -> async proxy
    // lane r owns peer r: its bulk copy, the wait for that copy's completion and the flag — the release store of the
    // SAME thread orders the flag behind the completed copy, so no warp-wide system fence is needed on this path
    if (lane < world && lane != me) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(xch_slot(x, seq, lane, me)),
                   "r"(smem_u32(stage)), "r"((unsigned)(n2 * 8)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(xch_flag(x, seq, lane, me)), "r"(seq) : "memory");
    }
    __syncwarp();
  } else {
    const unsigned long long v = lane < n ? __ldcg(local + lane) : 0ull;
    for (int r = 0; r < world; ++r) {
      if (r == me || lane >= n) continue;
      reinterpret_cast<volatile unsigned long long*>(xch_slot(x, seq, r, me))[lane] = v;
    }
    __threadfence_system();
    __syncwarp();
    if (lane < world && lane != me)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(xch_flag(x, seq, lane, me)), "r"(seq) : "memory");
  }
  if (lane == 0) stamps[2] = gtimer();
}

// One warp: wait until every peer's block of step `seq` has landed in this rank's mailbox (readable with volatile loads
// at xch_slot(x, seq, x.rank, r)).  false = a peer timed out (*status = 1).
static __device__ __noinline__ bool xch_wait(const XchCtx& x, unsigned seq, int lane) {
  const int world = x.world, me = x.rank;
  unsigned long long* stamps = xch_stamps(x, seq);
  if (lane == 0) stamps[4] = gtimer();
  bool ok = true;
  if (lane < world && lane != me) {
    const unsigned int* f = xch_flag(x, seq, me, lane);
    const unsigned long long t0 = gtimer();
    const unsigned long long limit = (unsigned long long)(x.timeout_ms ? x.timeout_ms : 2000u) * 1000000ull;
    for (;;) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v == seq) break;
      if (gtimer() - t0 > limit) { ok = false; break; }
      __nanosleep(100);
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  __threadfence_system();
  if (lane == 0) stamps[3] = gtimer();
  if (!ok && lane == 0 && x.status) *x.status = 1;
  return ok;
}

static __device__ __forceinline__ bool xch_publish_and_wait(const XchCtx& x, unsigned seq, const unsigned long long* local, int n,
                                                            int lane, unsigned long long* stage) {
  xch_publish(x, seq, local, n, lane, stage);
  return xch_wait(x, seq, lane);
}

// totals[e] += sum over ranks of step `seq`'s block, which this rank has published before; the local block is left
// zero.  One warp.  `scratch`: 16-byte aligned shared memory of xch_scratch_bytes(world, n) bytes.
// Every source — the local block and each peer's copy in this rank's mailbox — comes in as ONE bulk copy (TMA, global ->
// shared), all in flight together behind one mbarrier, and the running totals as one batch of register loads: a single
// loaded-L2 round trip instead of one per word (under the streaming kernels' traffic a dependent load costs 1-3 us,
// and whoever runs this holds an SM).
__host__ __device__ inline size_t xch_scratch_bytes(int world, int n) { return 16 + (size_t)(world > 0 ? world : 1) * (((size_t)n + 1) & ~(size_t)1) * 8; }

static __device__ __noinline__ void xch_consume_block_i64(const XchCtx& x, unsigned seq, unsigned long long* block, int n,
                                                          long long* totals, int lane, unsigned char* scratch) {
  const bool ok = x.world > 1 ? xch_wait(x, seq, lane) : true;
  const int n2 = (n + 1) & ~1;
  const uint32_t bytes = (uint32_t)n2 * 8u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(scratch);
  unsigned long long* data = reinterpret_cast<unsigned long long*>(scratch + 16);
  const int nsrc = ok ? x.world : 1;                          // source 0: the local block; then the peers in rank order
  if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncwarp();
  if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)nsrc * bytes);
  __syncwarp();
  if (lane < nsrc) {
    const void* src = block;
    if (lane > 0) { const int r = lane - 1 < x.rank ? lane - 1 : lane; src = xch_slot(x, seq, x.rank, r); }
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(data + (size_t)lane * n2)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
  }
  for (int e0 = lane; e0 < n; e0 += 32 * 16) {
    long long tot[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { const int e = e0 + 32 * u; tot[u] = e < n ? __ldcg(totals + e) : 0ll; }
    if (e0 == lane) mbar_wait(bar, 0u);                       // the bulk copies land while the totals are in flight
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = e0 + 32 * u;
      if (e < n) {
        long long sum = 0;
        for (int sidx = 0; sidx < nsrc; ++sidx) sum += (long long)data[(size_t)sidx * n2 + e];
        totals[e] = tot[u] + sum;
        block[e] = 0ull;
      }
    }
  }
  __syncwarp();
}

}  // namespace lhn
