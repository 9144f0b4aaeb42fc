// Diagnostic: does a tiled tensor-map copy (cp.async.bulk.tensor.4d) of FP64 accept an ODD or NEGATIVE innermost start coordinate?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_coord_test scripts/diag/tma_coord_test.cu && gpurun -- gpurun_out/tma_coord_test
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include "../../quokka_b200/csrc/qk_tma.cuh"

struct alignas(64) Map {
	unsigned char b[128];
};
__global__ void k(const __grid_constant__ Map m, int x, int y, int z, double *out, int n)
{
	extern __shared__ __align__(128) unsigned char raw[];
	double *s = reinterpret_cast<double *>(raw);
	uint64_t *bar = reinterpret_cast<uint64_t *>(raw + ((n * 8 + 127) / 128) * 128);
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		mbar_init_fence();
	}
	__syncwarp();
	if (elect_one()) {
		mbar_arrive_expect_tx(bar, (unsigned)n * 8u);
		tma_tile_g2s(s, &m, x, y, z, 0, bar);
	}
	mbar_wait(bar, 0);
	for (int i = threadIdx.x; i < n; i += 32)
		out[i] = s[i];
}
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
			   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
	const int NX = 48, NY = 8, NZ = 8, NC = 7, BW = 38;
	std::vector<double> h((size_t)NX * NY * NZ * NC);
	for (size_t i = 0; i < h.size(); ++i)
		h[i] = (double)i;
	double *d, *o;
	cudaMalloc(&d, h.size() * 8);
	cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
	const int n = BW * NC;
	cudaMalloc(&o, n * 8);
	void *p = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
	enc_fn enc = (enc_fn)p;
	Map m;
	const cuuint64_t dims[4] = {NX, NY, NZ, NC};
	const cuuint64_t strides[3] = {NX * 8, NX * NY * 8, NX * NY * NZ * 8};
	const cuuint32_t box[4] = {BW, 1, 1, NC}, es[4] = {1, 1, 1, 1};
	CUresult r = enc((CUtensorMap *)&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
			 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode %d\n", (int)r);
	const int xs[] = {0, 2, 1, 3, -2, -1, 11, 20, 47};
	for (int x : xs) {
		cudaMemset(o, 0xff, n * 8);
		k<<<1, 32, ((n * 8 + 127) / 128) * 128 + 64>>>(m, x, 3, 2, o, n);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) {
			printf("x = %d: %s\n", x, cudaGetErrorString(e));
			return 1;
		}
		std::vector<double> g(n);
		cudaMemcpy(g.data(), o, n * 8, cudaMemcpyDeviceToHost);
		int bad = 0;
		for (int c = 0; c < NC; ++c)
			for (int i = 0; i < BW; ++i) {
				const int xi = x + i;
				const double want = (xi < 0 || xi >= NX) ? 0.0 : h[(size_t)xi + (size_t)NX * (3 + NY * (2 + (size_t)NZ * c))];
				if (g[c * BW + i] != want)
					++bad;
			}
		printf("x = %d: ok launch, %d mismatches\n", x, bad);
	}
	return 0;
}
