// qk_tma.cuh -- sm_100a async-copy primitives used by the staged sweep kernels: 1-D bulk copies by the TMA engine
// (cp.async.bulk, SASS UBLKCP) completing on shared-memory mbarriers.  One elected lane of a warp arms the barrier
// with the byte count and issues the copies; every lane of that warp waits on the barrier's phase parity.
#pragma once
#include <stdint.h>

// One elected lane of a converged warp (elect.sync): unlike `lane == 0`, the predicate tells the compiler that exactly one lane
// takes the branch, so the bulk copies inside it are issued from the uniform datapath without a per-lane waterfall loop.
__device__ __forceinline__ bool elect_one()
{
	unsigned pred;
	asm volatile("{\n"
		     ".reg .pred p;\n"
		     "elect.sync _|p, 0xffffffff;\n"
		     "selp.u32 %0, 1, 0, p;\n"
		     "}\n"
		     : "=r"(pred));
	return pred != 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	asm volatile("{\n"
		     ".reg .pred p;\n"
		     "QK_WAIT_%=:\n"
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		     "@p bra QK_DONE_%=;\n"
		     "bra QK_WAIT_%=;\n"
		     "QK_DONE_%=:\n"
		     "}\n" ::"r"(smem_u32(bar)),
		     "r"(parity)
		     : "memory");
}

// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
		     "r"(smem_u32(bar))
		     : "memory");
}

// ---- tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG): one copy moves a [x-tile, rows, components] box of a 4-D array (x, y, z, component)
// described by a CUtensorMap the host encoded (cuTensorMapEncodeTiled); elements outside the array are zero-filled and still counted in the
// transaction bytes, so ragged tiles need no special cases.  dst must be 128-byte aligned.
__device__ __forceinline__ void tma_tile_g2s(void *dst, const void *tmap, int x, int y, int z, int n, uint64_t *bar)
{
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
		     "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(n), "r"(smem_u32(bar))
		     : "memory");
}
constexpr int TMAP_BYTES = 128; // sizeof(CUtensorMap)
constexpr int TMAP_MAXB = 24;	// boxes per launch of a kernel that carries its descriptors as __grid_constant__ parameters
struct alignas(64) TmapBytes {
	unsigned char b[TMAP_BYTES];
};
// for descriptors that live in global memory and are rewritten between launches: acquire them before the first use in a warp (unused: the
// sweep kernels take their descriptors as parameters, where no fence is needed -- measured: the fence cost the x sweep 27 %)
__device__ __forceinline__ void tmap_acquire(const void *tmap)
{
	asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
