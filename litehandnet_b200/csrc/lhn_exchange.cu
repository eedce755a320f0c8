// lhn_exchange_flush: the exchange of the LAST step's block (the in-kernel exchange of lhn_decode_heatmap_pck_xch runs one
// launch behind, so the final block of a sequence is still local when the sequence ends).  One warp.
#include "lhn_exchange.cuh"

namespace lhn {

int check_launch();

__global__ void __launch_bounds__(32) xch_flush_kernel(const XchCtx x, unsigned seq, unsigned long long* block, int n,
                                                       long long* totals) {
  extern __shared__ __align__(16) unsigned long long stage[];
  xch_allreduce_block_i64(x, seq, block, n, totals, threadIdx.x, stage);
}

}  // namespace lhn

using namespace lhn;

extern "C" int lhn_exchange_flush(const lhn_exchange* xch, int64_t* block, int n, int64_t* totals, lhn_stream_t stream) {
  if (!xch || !block || !totals || n <= 0 || (int64_t)n * 8 > LHN_XCH_PAYLOAD_BYTES) return LHN_EINVAL;
  if (xch->world < 1 || xch->world > LHN_XCH_MAX_RANKS || xch->rank < 0 || xch->rank >= xch->world || xch->seq == 0) return LHN_EINVAL;
  XchCtx x{};
  for (int r = 0; r < xch->world; ++r) {
    if (!xch->mailbox[r]) return LHN_EINVAL;
    x.mail[r] = static_cast<unsigned char*>(xch->mailbox[r]);
  }
  x.world = xch->world; x.rank = xch->rank; x.timeout_ms = xch->timeout_ms; x.status = xch->status;
  xch_flush_kernel<<<1, 32, (size_t)(n + 2) * 8, (cudaStream_t)stream>>>(x, xch->seq, reinterpret_cast<unsigned long long*>(block), n,
                                                                        reinterpret_cast<long long*>(totals));
  return check_launch();
}
