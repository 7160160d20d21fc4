// SASS probe: one Montgomery multiplication / one squaring per kernel, so that `cuobjdump -sass` shows the instruction mix
// of Fp::mul and Fp::sqr in isolation (profiles/r2_sass_summary.txt).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -cubin
#include "field.cuh"
using namespace h2v;
__global__ void probe_fq_mul(const Fq* a, const Fq* b, Fq* o) { o[threadIdx.x] = Fq::mul(a[threadIdx.x], b[threadIdx.x]); }
__global__ void probe_fq_sqr(const Fq* a, Fq* o) { o[threadIdx.x] = a[threadIdx.x].sqr(); }
__global__ void probe_fr_mul(const Fr* a, const Fr* b, Fr* o) { o[threadIdx.x] = Fr::mul(a[threadIdx.x], b[threadIdx.x]); }
__global__ void probe_fq_mul_lohi(const Fq* a, const Fq* b, Fq* o) { o[threadIdx.x] = Fq::mul_any(a[threadIdx.x], b[threadIdx.x]); }
