// lhn_exchange_flush: completes the exchange of the last steps (the in-kernel exchange of lhn_decode_heatmap_pck_xch is
// pipelined over two launches, so when a sequence ends the last block is still local and the one before it is published
// but not yet added).  One warp.
#include "lhn_exchange.cuh"

namespace lhn {

int check_launch();

__global__ void __launch_bounds__(32) xch_flush_kernel(const XchCtx x, unsigned seq2, unsigned long long* block2, unsigned seq,
                                                       unsigned long long* block, int n, long long* totals) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x;
  unsigned char* scratch = smem + (((size_t)n + 2) * 8 + 15) / 16 * 16;       // after the publish stage
  if (block2) xch_consume_block_i64(x, seq2, block2, n, totals, lane, scratch);   // published by the last launch
  if (block) {
    if (x.world > 1) xch_publish(x, seq, block, n, lane, reinterpret_cast<unsigned long long*>(smem));   // the last launch's own block
    xch_consume_block_i64(x, seq, block, n, totals, lane, scratch);
  }
}

}  // namespace lhn

using namespace lhn;

extern "C" int lhn_exchange_flush(const lhn_exchange* xch, int n, int64_t* totals, lhn_stream_t stream) {
  if (!xch || !totals || n <= 0 || (int64_t)n * 8 > LHN_XCH_PAYLOAD_BYTES) return LHN_EINVAL;
  if (xch->world < 1 || xch->world > LHN_XCH_MAX_RANKS || xch->rank < 0 || xch->rank >= xch->world) return LHN_EINVAL;
  if ((xch->prev_block && xch->prev_seq == 0) || (xch->prev2_block && xch->prev2_seq == 0)) return LHN_EINVAL;
  if (!xch->prev_block && !xch->prev2_block) return LHN_OK;
  XchCtx x{};
  for (int r = 0; r < xch->world; ++r) {
    if (!xch->mailbox[r]) return LHN_EINVAL;
    x.mail[r] = static_cast<unsigned char*>(xch->mailbox[r]);
  }
  x.world = xch->world; x.rank = xch->rank; x.timeout_ms = xch->timeout_ms; x.status = xch->status;
  const size_t smem = (((size_t)n + 2) * 8 + 15) / 16 * 16 + xch_scratch_bytes(x.world, n);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(xch_flush_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return LHN_ECUDA;
  }
  xch_flush_kernel<<<1, 32, smem, (cudaStream_t)stream>>>(
      x, xch->prev2_seq, static_cast<unsigned long long*>(xch->prev2_block), xch->prev_seq,
      static_cast<unsigned long long*>(xch->prev_block), n, reinterpret_cast<long long*>(totals));
  return check_launch();
}
