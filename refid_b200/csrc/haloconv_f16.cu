// fp16-storage instantiations of the halo-conv kernel (forward-only / inference plans); see haloconv_kernel.cuh.
#include "haloconv_kernel.cuh"

namespace refid {

int launch_haloconv_f16(const HaloConvParams& p, int BN, int NM, cudaStream_t stream) {
  return launch_haloconv_flavour<true>(p, BN, NM, stream);
}

}  // namespace refid
