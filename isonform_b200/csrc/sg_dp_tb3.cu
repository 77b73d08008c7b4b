// sg_dp_tb3.cu -- the DP kernel under tie-break variant 3 (see sg_dp.cuh); a translation unit of its own so
// that the variants compile in parallel.
#include "sg_dp.cuh"

void isof_launch_dp_tb3(int blocks, cudaStream_t st, const AlignParams &A, int64_t start, int64_t end, unsigned long long *counter)
{
    launch_sg_dp<3>(blocks, st, A, start, end, counter);
}

unsigned isof_viol_dp_tb3() { return isof_read_violations_tu(); }
