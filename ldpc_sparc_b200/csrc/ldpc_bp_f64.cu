// fp64 instantiation of the BP kernels (the parity mode and the legacy ABI).
#include "ldpc_bp_kernels.cuh"
namespace lsb_bp {
template int launch_qc<double>(const QcLaunch<double> &);
template int launch_generic<double>(const GenericLaunch<double> &);
template int generic_blocks_per_sm<double>();

__global__ void lxor_kernel(double a, double b, int corr, double *out)
{
    *out = corr ? lxor<double, true>(a, b) : lxor<double, false>(a, b);
}

__global__ void lxfb_kernel(double *L, int dc, int corr, double *out)
{
    *out = corr ? check_node_dyn<double, LSB_BP_SUMPROD2>(L, 1, dc, 1.0)
                : check_node_dyn<double, LSB_BP_MINSUM>(L, 1, dc, 1.0);
}

void launch_lxor(double a, double b, int corr, double *d_out)
{
    lxor_kernel<<<1, 1>>>(a, b, corr, d_out);
    lsb::count_launch();
}

void launch_lxfb(double *d_L, int dc, int corr, double *d_out)
{
    lxfb_kernel<<<1, 1>>>(d_L, dc, corr, d_out);
    lsb::count_launch();
}
}  // namespace lsb_bp
