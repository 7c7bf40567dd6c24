// fp32 instantiation of the BP kernels (the throughput path).
#include "ldpc_bp_kernels.cuh"
namespace lsb_bp {
template int launch_qc<float>(const QcLaunch<float> &);
template int launch_generic<float>(const GenericLaunch<float> &);
template int generic_blocks_per_sm<float>();
}  // namespace lsb_bp
