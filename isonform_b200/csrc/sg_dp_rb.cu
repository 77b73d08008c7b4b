// sg_dp_rb.cu -- the DP kernel of the re-based 16-bit class (pairs beyond the plain 16-bit score range, e.g. 10 kb
// reads; see dp_pair16<., RB> in sg_dp.cuh), tie-break variants 0 and 1 (2 and 3: sg_dp_rb2.cu); translation units of
// their own so that they compile in parallel with the others.
#include "sg_dp.cuh"

void isof_launch_dp_rb23(int tb_variant, int blocks, cudaStream_t st, const AlignParams &A, int64_t start, unsigned long long *counter);

void isof_launch_dp_rb(int tb_variant, int blocks, cudaStream_t st, const AlignParams &A, int64_t start, unsigned long long *counter)
{
    switch (tb_variant) {
        case 0: sg_dp_rb_kernel<0><<<blocks, DP_WARPS * 32, 0, st>>>(A, start, counter); break;
        case 1: sg_dp_rb_kernel<1><<<blocks, DP_WARPS * 32, 0, st>>>(A, start, counter); break;
        default: isof_launch_dp_rb23(tb_variant, blocks, st, A, start, counter); break;
    }
}

unsigned isof_viol_dp_rb23();
unsigned isof_viol_dp_rb() { return isof_read_violations_tu() | isof_viol_dp_rb23(); }
