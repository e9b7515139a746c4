// blind_rotate_exact.cu -- placeholder until the exact-mode kernel lands.
#include "kernels.cuh"
namespace tfhe_b200 {
cudaError_t launch_blind_rotate_exact(const BrArgs &, const double *, cudaStream_t, uint64_t *) { return cudaErrorNotSupported; }
}  // namespace tfhe_b200
