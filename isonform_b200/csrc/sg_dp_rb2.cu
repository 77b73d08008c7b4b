// sg_dp_rb2.cu -- tie-break variants 2 and 3 of the re-based 16-bit DP kernel (see sg_dp_rb.cu).
#include "sg_dp.cuh"

void isof_launch_dp_rb23(int tb_variant, int blocks, cudaStream_t st, const AlignParams &A, int64_t start, unsigned long long *counter)
{
    if (tb_variant == 2) sg_dp_rb_kernel<2><<<blocks, DP_WARPS * 32, 0, st>>>(A, start, counter);
    else sg_dp_rb_kernel<3><<<blocks, DP_WARPS * 32, 0, st>>>(A, start, counter);
}

unsigned isof_viol_dp_rb23() { return isof_read_violations_tu(); }
