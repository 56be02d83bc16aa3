// riemann_probe.cu -- instruction-budget probe: one Riemann solve / one PPM cell per thread,
// compiled standalone so the SASS of the building blocks can be counted without the kernel
// bodies around them (scripts/sass_count.py).
#include "../../artemis_b200/csrc/march.cuh"
using namespace ab200;

template <int RS, int FLUID>
__global__ void k_probe_riemann(const double *__restrict__ in, double *__restrict__ out, EosConsts eos, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double wl[6], wr[6], o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int v = 0; v < 6; ++v) { wl[v] = in[(2 * v) * n + t]; wr[v] = in[(2 * v + 1) * n + t]; }
  Riemann<RS, FLUID>::solve(eos, wl, wr, o);
#pragma unroll
  for (int v = 0; v < 8; ++v) out[v * n + t] = o[v];
}
template __global__ void k_probe_riemann<AB200_HLLC, AB200_GAS>(const double *, double *, EosConsts, int);
template __global__ void k_probe_riemann<AB200_HLLE, AB200_GAS>(const double *, double *, EosConsts, int);
template __global__ void k_probe_riemann<AB200_LLF, AB200_GAS>(const double *, double *, EosConsts, int);
template __global__ void k_probe_riemann<AB200_HLLE, AB200_DUST>(const double *, double *, EosConsts, int);

// 6 variables: interface value + monotonisation of one cell
__global__ void k_probe_ppm(const double *__restrict__ in, double *__restrict__ out, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int v = 0; v < 6; ++v) {
    const double *q = in + (size_t)v * 5 * n + t;
    const double ilo = ppm_iface(q[0], q[n], q[2 * n], q[3 * n]);
    const double iup = ppm_iface(q[n], q[2 * n], q[3 * n], q[4 * n]);
    double ql, qr;
    ppm_mono(ilo, q[2 * n], iup, ql, qr);
    out[(2 * v) * n + t] = ql;
    out[(2 * v + 1) * n + t] = qr;
  }
}
__global__ void k_probe_plm(const double *__restrict__ in, double *__restrict__ out, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int v = 0; v < 6; ++v) {
    const double *q = in + (size_t)v * 3 * n + t;
    double ql, qr;
    plm(q[0], q[n], q[2 * n], ql, qr);
    out[(2 * v) * n + t] = ql;
    out[(2 * v + 1) * n + t] = qr;
  }
}
