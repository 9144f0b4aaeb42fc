// diagnostic (not part of the library): exact vs relaxed HLL face flux of the radiation sweeps on random admissible and inadmissible states
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=true -std=c++17 -I quokka_b200/csrc -I include scripts/diag/radflux_ab.cu -o /tmp/radflux_ab
#include "qk_rad_kernels.cuh"
#include <cstdio>
#include <vector>
struct Case {
	double L[4], R[4], cL[4], cR[4];
};
__global__ void k(RadConst c, const Case *cs, int n, double *out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	double Fe[4], Fr[4];
	// conserved states either side in a 4-component array with stride 1
	rad_face_flux<0>(c, cs[i].L, cs[i].R, cs[i].cL, cs[i].cR, 1, Fe);
	rad_face_flux_r<0>(c, cs[i].L, cs[i].R, cs[i].cL, cs[i].cR, 1, Fr);
	for (int m = 0; m < 4; ++m) {
		out[8 * i + m] = Fe[m];
		out[8 * i + 4 + m] = Fr[m];
	}
}
int qk_require_device() { return 0; }
int64_t g_qk_launches = 0;
static double rnd(unsigned long long &s)
{
	s ^= s << 13;
	s ^= s >> 7;
	s ^= s << 17;
	return (double)(s >> 11) * (1.0 / 9007199254740992.0);
}
int main()
{
	qk_rad_params p{};
	p.c_light = 2.99792458e10;
	p.c_hat = p.c_light / 30.0;
	p.Erad_floor = 1e-12;
	p.ngroups = 1;
	p.nstart = 6;
	RadConst c = make_rad_const(&p);
	const int n = 1 << 16;
	std::vector<Case> h(n);
	unsigned long long s = 88172645463325252ull;
	for (int i = 0; i < n; ++i) {
		for (int side = 0; side < 2; ++side) {
			double *q = side ? h[i].R : h[i].L, *cc = side ? h[i].cR : h[i].cL;
			int kind = (int)(rnd(s) * 6);
			double E = pow(10.0, 4 * rnd(s) - 2), f = 0.95 * rnd(s);
			if (kind == 1) E = -E;
			if (kind == 2) f = 1.3;
			if (kind == 3) f = 1.0 - 1e-12;
			if (kind == 4) f = 0.0;
			double mu = 2 * rnd(s) - 1, ph = 6.283185307179586 * rnd(s), st = sqrt(1 - mu * mu);
			q[0] = E; q[1] = f * st * cos(ph); q[2] = f * st * sin(ph); q[3] = f * mu;
			// conserved state of the adjacent cell: another random state
			double E2 = pow(10.0, 4 * rnd(s) - 2) * ((rnd(s) < 0.15) ? -1 : 1), f2 = (rnd(s) < 0.2) ? 1.3 : 0.9 * rnd(s);
			cc[0] = E2; cc[1] = f2 * 0.6 * p.c_light * fabs(E2); cc[2] = f2 * 0.64 * p.c_light * fabs(E2); cc[3] = -f2 * 0.48 * p.c_light * fabs(E2);
		}
	}
	Case *d; double *o;
	cudaMalloc(&d, sizeof(Case) * n); cudaMalloc(&o, 64 * n);
	cudaMemcpy(d, h.data(), sizeof(Case) * n, cudaMemcpyHostToDevice);
	k<<<n / 256, 256>>>(c, d, n, o);
	std::vector<double> r(8 * n);
	cudaMemcpy(r.data(), o, 64 * n, cudaMemcpyDeviceToHost);
	int nbad = 0;
	for (int i = 0; i < n; ++i) {
		double worst = 0;
		for (int m = 0; m < 4; ++m) {
			double a = r[8 * i + m], b = r[8 * i + 4 + m];
			double sc = fabs(a) + 1e-300;
			if (std::isnan(a) != std::isnan(b)) worst = 1e99;
			else if (!std::isnan(a)) worst = fmax(worst, fabs(a - b) / sc);
		}
		if (worst > 1e-10) {
			if (nbad < 12) {
				printf("case %d worst %.3e\n  L = %.17g %.17g %.17g %.17g\n  R = %.17g %.17g %.17g %.17g\n  cL = %.6g %.6g %.6g %.6g  cR = %.6g %.6g %.6g %.6g\n  exact   %.10g %.10g %.10g %.10g\n  relaxed %.10g %.10g %.10g %.10g\n",
				       i, worst, h[i].L[0], h[i].L[1], h[i].L[2], h[i].L[3], h[i].R[0], h[i].R[1], h[i].R[2], h[i].R[3], h[i].cL[0], h[i].cL[1], h[i].cL[2],
				       h[i].cL[3], h[i].cR[0], h[i].cR[1], h[i].cR[2], h[i].cR[3], r[8 * i], r[8 * i + 1], r[8 * i + 2], r[8 * i + 3], r[8 * i + 4], r[8 * i + 5],
				       r[8 * i + 6], r[8 * i + 7]);
			}
			++nbad;
		}
	}
	printf("%d of %d cases differ by more than 1e-10\n", nbad, n);
	return 0;
}
