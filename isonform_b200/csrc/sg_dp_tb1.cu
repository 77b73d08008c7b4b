// sg_dp_tb1.cu -- the DP kernel under tie-break variant 1 (see sg_dp.cuh); a translation unit of its own so
// that the variants compile in parallel.
#include "sg_dp.cuh"

void isof_launch_dp_tb1(int blocks, cudaStream_t st, const AlignParams &A, int64_t start, int64_t end, unsigned long long *counter)
{
    launch_sg_dp<1>(blocks, st, A, start, end, counter);
}

unsigned isof_viol_dp_tb1() { return isof_read_violations_tu(); }
