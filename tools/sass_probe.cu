#include "field.cuh"
using namespace zkb;
__global__ void probe_fq_mul(const uint4* a, const uint4* b, uint4* o) {
    Fq x = Fq::load(a + 2 * threadIdx.x), y = Fq::load(b + 2 * threadIdx.x);
    fp_mul_lazy(x, y).store(o + 2 * threadIdx.x);
}
