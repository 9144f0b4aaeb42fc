// qk_comm.cpp -- rank<->rank transport of the hot path: NCCL over NVLink/NVSwitch replacing the MPI
// point-to-point + Allreduce calls AMReX makes for FillBoundary and the scalar reductions
// (extern/amrex/Src/Base/AMReX_FabArrayCommI.H:7-165; AMReX_ParallelDescriptor.cpp:1091,1659,1746).
// One process per GPU.  NCCL is resolved at run time (dlopen), preferring a copy that the host
// application has already loaded, so libquokka_b200.so has no link-time dependency on it.
#include "qk_level.h"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

namespace
{
struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	bool ok = false;
} g_nccl;

int load_nccl()
{
	if (g_nccl.ok)
		return 0;
	const char *names[] = {"libnccl.so.2", "libnccl.so"};
	void *h = nullptr;
	for (const char *n : names) {
		h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); // already in the process (e.g. torch's)?
		if (h)
			break;
	}
	for (int i = 0; i < 2 && !h; ++i)
		h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
	if (!h)
		return QK_ERR_UNSUPPORTED;
	g_nccl.h = h;
#define QK_SYM(field, name)                                                                                                                          \
	*(void **)(&g_nccl.field) = dlsym(h, name);                                                                                                  \
	if (!g_nccl.field)                                                                                                                           \
		return QK_ERR_UNSUPPORTED;
	QK_SYM(GetUniqueId, "ncclGetUniqueId");
	QK_SYM(CommInitRank, "ncclCommInitRank");
	QK_SYM(CommDestroy, "ncclCommDestroy");
	QK_SYM(Send, "ncclSend");
	QK_SYM(Recv, "ncclRecv");
	QK_SYM(AllReduce, "ncclAllReduce");
	QK_SYM(GroupStart, "ncclGroupStart");
	QK_SYM(GroupEnd, "ncclGroupEnd");
#undef QK_SYM
	g_nccl.ok = true;
	return 0;
}
inline int nc(ncclResult_t r) { return r == ncclSuccess ? 0 : (1000 + (int)r); }
} // namespace

struct qk_comm {
	ncclComm_t comm = nullptr;
	int rank = 0, nranks = 1;
	void *d_scalar = nullptr;
	void *h_scalar = nullptr;
};

extern "C" int qk_comm_unique_id(void *id128)
{
	if (!id128)
		return QK_ERR_BAD_ARG;
	int rc = load_nccl();
	if (rc)
		return rc;
	ncclUniqueId id;
	rc = nc(g_nccl.GetUniqueId(&id));
	if (rc)
		return rc;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	memcpy(id128, &id, 128);
	return 0;
}

extern "C" int qk_comm_create(const void *id128, int rank, int nranks, qk_comm **out)
{
	if (!id128 || !out || rank < 0 || rank >= nranks)
		return QK_ERR_BAD_ARG;
	int rc = qk_require_device();
	if (rc)
		return rc;
	rc = load_nccl();
	if (rc)
		return rc;
	qk_comm *c = new qk_comm();
	c->rank = rank;
	c->nranks = nranks;
	ncclUniqueId id;
	memcpy(&id, id128, 128);
	rc = nc(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
	if (rc) {
		delete c;
		return rc;
	}
	if (cudaMalloc(&c->d_scalar, 64) != cudaSuccess || cudaMallocHost(&c->h_scalar, 64) != cudaSuccess) {
		delete c;
		return QK_ERR_NOMEM;
	}
	*out = c;
	return 0;
}

extern "C" void qk_comm_destroy(qk_comm *c)
{
	if (!c)
		return;
	if (c->comm)
		g_nccl.CommDestroy(c->comm);
	if (c->d_scalar)
		cudaFree(c->d_scalar);
	if (c->h_scalar)
		cudaFreeHost(c->h_scalar);
	delete c;
}

extern "C" int qk_comm_rank(const qk_comm *c) { return c ? c->rank : 0; }
extern "C" int qk_comm_nranks(const qk_comm *c) { return c ? c->nranks : 1; }

extern "C" int qk_comm_group_start(qk_comm *) { return nc(g_nccl.GroupStart()); }
extern "C" int qk_comm_group_end(qk_comm *) { return nc(g_nccl.GroupEnd()); }
extern "C" int qk_comm_send(qk_comm *c, const void *buf, size_t bytes, int peer, cudaStream_t s)
{
	return nc(g_nccl.Send(buf, bytes, ncclInt8, peer, c->comm, s));
}
extern "C" int qk_comm_recv(qk_comm *c, void *buf, size_t bytes, int peer, cudaStream_t s)
{
	return nc(g_nccl.Recv(buf, bytes, ncclInt8, peer, c->comm, s));
}

extern "C" int qk_comm_allreduce_sum_i64(qk_comm *c, int64_t *v, cudaStream_t s)
{
	if (!c || c->nranks == 1)
		return 0;
	memcpy(c->h_scalar, v, 8);
	QK_CUDA(cudaMemcpyAsync(c->d_scalar, c->h_scalar, 8, cudaMemcpyHostToDevice, s));
	int rc = nc(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclInt64, ncclSum, c->comm, s));
	if (rc)
		return rc;
	QK_CUDA(cudaMemcpyAsync(c->h_scalar, c->d_scalar, 8, cudaMemcpyDeviceToHost, s));
	QK_CUDA(cudaStreamSynchronize(s));
	memcpy(v, c->h_scalar, 8);
	return 0;
}

// in-place all-reduce of `count` unsigned 64-bit values that already live in DEVICE memory (sum, or max when is_max): no staging copies and no
// host synchronisation -- the caller reads the result with the D2H copy it makes anyway (cell counters of a stage; the order-preserving keys
// of the signal-speed maxima)
extern "C" int qk_comm_allreduce_dev_u64(qk_comm *c, unsigned long long *d_vals, int count, int is_max, cudaStream_t s)
{
	if (!c || c->nranks == 1)
		return 0;
	return nc(g_nccl.AllReduce(d_vals, d_vals, (size_t)count, ncclUint64, is_max ? ncclMax : ncclSum, c->comm, s));
}

extern "C" int qk_comm_allreduce_max_f64(qk_comm *c, double *v, cudaStream_t s)
{
	if (!c || c->nranks == 1)
		return 0;
	memcpy(c->h_scalar, v, 8);
	QK_CUDA(cudaMemcpyAsync(c->d_scalar, c->h_scalar, 8, cudaMemcpyHostToDevice, s));
	int rc = nc(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclFloat64, ncclMax, c->comm, s));
	if (rc)
		return rc;
	QK_CUDA(cudaMemcpyAsync(c->h_scalar, c->d_scalar, 8, cudaMemcpyDeviceToHost, s));
	QK_CUDA(cudaStreamSynchronize(s));
	memcpy(v, c->h_scalar, 8);
	return 0;
}
