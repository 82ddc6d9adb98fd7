// k0_parse.cu — K0 on the device: one CTA (one warp, lane 0 walks the syntax) per substream chain.
//
// CABAC is serial inside a substream, so the parallelism is across substreams: the CTB rows of WPP pictures
// (wavefront-synchronised with the row above, kernels/k0_core.cuh) and across the pictures / grid tiles / files of a
// batch — a 12 MP iPhone-style grid image is 48 tiles x 8 rows = 384 chains; a batch of 8 such files keeps 3072
// chains in flight. Chains are ordered row-major across all pictures so that a chain only ever waits for a chain
// with a smaller index (scheduled no later than itself). The context table of a chain lives in shared memory.
#include "launch.h"
#include "k0_core.cuh"

namespace hc {

__global__ void __launch_bounds__(32, 20) k0_parse_kernel(const k0::Tables* __restrict__ tables, const k0::Pic* __restrict__ pics,
                                                      const k0::Sub* __restrict__ subs, const k0::Chain* __restrict__ chains,
                                                      int nchains) {
  __shared__ uint8_t ctx[k0::CTX_BYTES];
  if (threadIdx.x != 0 || (int)blockIdx.x >= nchains) return;
  const k0::Chain ch = chains[blockIdx.x];
  k0::Parser ps;
  ps.T = tables;
  ps.cabac.T = tables;
  ps.ctx = ctx;
  ps.P = nullptr;
  ps.sh = nullptr;
  ps.run_chain(pics, subs, ch.first_sub, ch.nsubs);
}

void launch_k0(const k0::Tables* tables, const k0::Pic* pics, const k0::Sub* subs, const k0::Chain* chains, int nchains,
               cudaStream_t stream) {
  if (nchains <= 0) return;
  k0_parse_kernel<<<nchains, 32, 0, stream>>>(tables, pics, subs, chains, nchains);
}

}  // namespace hc
