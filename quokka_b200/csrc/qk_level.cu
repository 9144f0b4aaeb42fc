// qk_level.cu -- the level object: ghost-cell fill (FillBoundary + physical BCs), the exchange plan
// (copy tags), pack/unpack for remote neighbours, and the FAITHFUL stage path of
// QuokkaSimulation::advanceHydroAtLevel built from the per-operator kernels (fluxes materialised,
// FOFC fallback).  The tuned fused path (qk_sweep.cu) hands over to this file whenever a stage
// flags a cell (redoFlag), so the rare first-order-flux-correction case keeps the reference's
// exact semantics (src/QuokkaSimulation.hpp:1146-1184, 1233-1271).
#include "qk_level.h"

#include <cuda.h> // CUtensorMap and its enums only: the encoder is resolved at run time (no link against libcuda)
#include "qk_kernels.cuh"

#include <algorithm>
#include <string.h>

// ------------------------------------------------------------------------------------------------
// copy-tag planning (host only; runs without a GPU): FabArray::FillBoundary(periodicity) semantics --
// every ghost cell of box b that, after a periodic shift, lies in the VALID region of box s is copied
// from s; corners included (cross=false, src/simulation.hpp:1755;
// extern/amrex/Src/Base/AMReX_FabArrayBase.cpp FB::define_fb).
// ------------------------------------------------------------------------------------------------
static inline int blen(const qk_box &b, int d) { return b.hi[d] - b.lo[d] + 1; }
static inline qk_box bgrow(qk_box b, int n)
{
	for (int d = 0; d < 3; ++d) {
		b.lo[d] -= n;
		b.hi[d] += n;
	}
	return b;
}

void qk_plan_tags(const qk_level &L, int ng, std::vector<HostTag> &out)
{
	out.clear();
	const int nb = (int)L.boxes.size();
	const qk_box dom = L.domain;
	int smin[3], smax[3];
	// amrex::Periodicity::shiftIntVect(nghost): a periodic direction shorter than the ghost width (one-cell-thick quasi-1-D / 2-D
	// domains) needs ceil(ng / length) images on either side, not one
	for (int d = 0; d < 3; ++d) {
		const int nimg = L.periodic[d] ? (ng + blen(dom, d) - 1) / blen(dom, d) : 0;
		smin[d] = -std::max(nimg, L.periodic[d] ? 1 : 0);
		smax[d] = -smin[d];
	}
	for (int b = 0; b < nb; ++b) {
		const qk_box g = bgrow(L.boxes[b], ng);
		for (int s = 0; s < nb; ++s) {
			if (L.owner[b] != L.my_rank && L.owner[s] != L.my_rank)
				continue;
			for (int sz = smin[2]; sz <= smax[2]; ++sz)
				for (int sy = smin[1]; sy <= smax[1]; ++sy)
					for (int sx = smin[0]; sx <= smax[0]; ++sx) {
						if (s == b && sx == 0 && sy == 0 && sz == 0)
							continue;
						const int sh[3] = {sx * blen(dom, 0), sy * blen(dom, 1), sz * blen(dom, 2)};
						HostTag t;
						bool empty = false;
						for (int d = 0; d < 3; ++d) {
							t.dst_region.lo[d] = std::max(g.lo[d], L.boxes[s].lo[d] + sh[d]);
							t.dst_region.hi[d] = std::min(g.hi[d], L.boxes[s].hi[d] + sh[d]);
							if (t.dst_region.lo[d] > t.dst_region.hi[d])
								empty = true;
							t.shift[d] = sh[d];
						}
						if (empty)
							continue;
						t.src_box = s;
						t.dst_box = b;
						t.src_rank = L.owner[s];
						t.dst_rank = L.owner[b];
						t.ncells = (int64_t)blen(t.dst_region, 0) * blen(t.dst_region, 1) * blen(t.dst_region, 2);
						t.offset = 0;
						out.push_back(t);
					}
		}
	}
}

// physical-boundary regions of one box (disjoint): x slabs span the full grown y,z range; y slabs only
// x inside the domain; z slabs only x,y inside.  The value written is the composition of the per-axis
// maps of amrex::FilccCell (AMReX_FilCC_3D_C.H:38-41,66-75,...), which is what the faces -> edges ->
// corners passes of PhysBCFunct produce (AMReX_PhysBCFunct.H:406-470).
static void plan_bc(const qk_level &L, int ng, int local_b, const qk_box &vb, std::vector<HostBcTag> &out)
{
	const qk_box g = bgrow(vb, ng);
	const qk_box dom = L.domain;
	qk_box in = g; // part of g inside the domain along the axes processed so far
	for (int d = 0; d < 3; ++d) {
		if (L.periodic[d])
			continue;
		if (g.lo[d] < dom.lo[d]) {
			HostBcTag t;
			t.box = local_b;
			t.region = in;
			t.region.hi[d] = dom.lo[d] - 1;
			out.push_back(t);
		}
		if (g.hi[d] > dom.hi[d]) {
			HostBcTag t;
			t.box = local_b;
			t.region = in;
			t.region.lo[d] = dom.hi[d] + 1;
			out.push_back(t);
		}
		in.lo[d] = std::max(in.lo[d], dom.lo[d]);
		in.hi[d] = std::min(in.hi[d], dom.hi[d]);
	}
}

// ------------------------------------------------------------------------------------------------
// device kernels: tag copies and physical BCs, all local boxes in ONE launch each
// ------------------------------------------------------------------------------------------------
namespace
{
// mode 0: same-rank copy (FB_local_copy_gpu, AMReX_FBI.H:272); 1: pack (:730); 2: unpack (:790)
template <int MODE>
__global__ void __launch_bounds__(256) k_tags(const DevTag *__restrict__ tags, const A4 *__restrict__ arrs, int scomp, int ncomp, double *__restrict__ buf)
{
	const DevTag t = tags[blockIdx.y];
	const int64_t ncell = t.ncells;
	const int nx = t.n[0], ny = t.n[1];
	for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (int64_t)gridDim.x * blockDim.x) {
		const int64_t jk = c / nx;
		const int i = t.lo[0] + (int)(c - jk * nx);
		const int k = t.lo[2] + (int)(jk / ny);
		const int j = t.lo[1] + (int)(jk - (jk / ny) * ny);
		for (int n = 0; n < ncomp; ++n) {
			if (MODE == 0) {
				arrs[t.dst](i, j, k, scomp + n) = arrs[t.src](i - t.sh[0], j - t.sh[1], k - t.sh[2], scomp + n);
			} else if (MODE == 1) {
				buf[t.off * ncomp + n * ncell + c] = arrs[t.src](i - t.sh[0], j - t.sh[1], k - t.sh[2], scomp + n);
			} else {
				arrs[t.dst](i, j, k, scomp + n) = buf[t.off * ncomp + n * ncell + c];
			}
		}
	}
}

// same for the int redoFlag (redoFlag.FillBoundary, QuokkaSimulation.hpp:1157)
template <int MODE>
__global__ void __launch_bounds__(256) k_tags_int(const DevTag *__restrict__ tags, const IA4 *__restrict__ arrs, int32_t *__restrict__ buf)
{
	const DevTag t = tags[blockIdx.y];
	const int64_t ncell = t.ncells;
	const int nx = t.n[0], ny = t.n[1];
	for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (int64_t)gridDim.x * blockDim.x) {
		const int64_t jk = c / nx;
		const int i = t.lo[0] + (int)(c - jk * nx);
		const int k = t.lo[2] + (int)(jk / ny);
		const int j = t.lo[1] + (int)(jk - (jk / ny) * ny);
		if (MODE == 0)
			arrs[t.dst](i, j, k) = arrs[t.src](i - t.sh[0], j - t.sh[1], k - t.sh[2]);
		else if (MODE == 1)
			buf[t.off + c] = arrs[t.src](i - t.sh[0], j - t.sh[1], k - t.sh[2]);
		else
			arrs[t.dst](i, j, k) = buf[t.off + c];
	}
}

__global__ void __launch_bounds__(256) k_phys_bc(const DevBcTag *__restrict__ tags, const A4 *__restrict__ arrs, int scomp, int ncomp, Box3 dom,
						 const int32_t *__restrict__ bc_lo, const int32_t *__restrict__ bc_hi, int per0, int per1, int per2)
{
	const DevBcTag t = tags[blockIdx.y];
	const int64_t ncell = (int64_t)t.n[0] * t.n[1] * t.n[2];
	const int nx = t.n[0], ny = t.n[1];
	const int per[3] = {per0, per1, per2};
	const A4 a = arrs[t.box];
	for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (int64_t)gridDim.x * blockDim.x) {
		const int64_t jk = c / nx;
		int idx[3];
		idx[0] = t.lo[0] + (int)(c - jk * nx);
		idx[2] = t.lo[2] + (int)(jk / ny);
		idx[1] = t.lo[1] + (int)(jk - (jk / ny) * ny);
		for (int n = scomp; n < scomp + ncomp; ++n) {
			int src[3] = {idx[0], idx[1], idx[2]};
			bool neg = false, skip = false;
#pragma unroll
			for (int d = 0; d < 3; ++d) {
				if (per[d])
					continue;
				if (idx[d] < dom.lo[d]) {
					const int bc = bc_lo[n * 3 + d];
					if (bc == QK_BC_REFLECT_EVEN || bc == QK_BC_REFLECT_ODD) {
						src[d] = 2 * dom.lo[d] - idx[d] - 1;
						neg ^= (bc == QK_BC_REFLECT_ODD);
					} else if (bc == QK_BC_FOEXTRAP) {
						src[d] = dom.lo[d];
					} else {
						skip = true; // ext_dir / int_dir: left for the caller (user Dirichlet functor)
					}
				} else if (idx[d] > dom.hi[d]) {
					const int bc = bc_hi[n * 3 + d];
					if (bc == QK_BC_REFLECT_EVEN || bc == QK_BC_REFLECT_ODD) {
						src[d] = 2 * dom.hi[d] - idx[d] + 1;
						neg ^= (bc == QK_BC_REFLECT_ODD);
					} else if (bc == QK_BC_FOEXTRAP) {
						src[d] = dom.hi[d];
					} else {
						skip = true;
					}
				}
			}
			if (skip)
				continue;
			const double v = a(src[0], src[1], src[2], n);
			a(idx[0], idx[1], idx[2], n) = neg ? -v : v;
		}
	}
}
} // namespace

// ------------------------------------------------------------------------------------------------
// level object
// ------------------------------------------------------------------------------------------------
static DevTag to_dev(const HostTag &t, int dst_local, int src_local)
{
	DevTag d;
	d.dst = dst_local;
	d.src = src_local;
	for (int k = 0; k < 3; ++k) {
		d.lo[k] = t.dst_region.lo[k];
		d.n[k] = blen(t.dst_region, k);
		d.sh[k] = t.shift[k];
	}
	d.pad = 0;
	d.ncells = t.ncells;
	d.off = t.offset;
	return d;
}

template <class T> static int upload(const std::vector<T> &h, T **d)
{
	*d = nullptr;
	if (h.empty())
		return 0;
	QK_CUDA(cudaMalloc(d, sizeof(T) * h.size()));
	QK_CUDA(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
	return 0;
}

int qk_exchange_plan::build(const qk_level &L, int ng, bool need_device)
{
	nghost = ng;
	send_cells_all = recv_cells_all = 0;
	std::vector<HostTag> all;
	qk_plan_tags(L, ng, all);
	local.clear();
	remote.clear();
	peers.clear();
	for (const HostTag &t : all) {
		if (t.src_rank == L.my_rank && t.dst_rank == L.my_rank)
			local.push_back(t);
		else
			remote.push_back(t);
	}
	// per-peer message layout: tags in plan order (dst box, src box, shift) -- both sides enumerate the
	// same global list, so offsets agree without communication
	std::vector<int> pr;
	for (const HostTag &t : remote)
		pr.push_back(t.src_rank == L.my_rank ? t.dst_rank : t.src_rank);
	std::vector<int> uniq = pr;
	std::sort(uniq.begin(), uniq.end());
	uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
	for (int p : uniq) {
		PeerPlan pp;
		pp.peer = p;
		pp.send_cells = pp.recv_cells = 0;
		for (size_t i = 0; i < remote.size(); ++i) {
			if (pr[i] != p)
				continue;
			HostTag &t = remote[i];
			if (t.src_rank == L.my_rank) {
				t.offset = pp.send_cells;
				pp.send_cells += t.ncells;
				pp.send.push_back(to_dev(t, -1, L.local_of[t.src_box]));
			} else {
				t.offset = pp.recv_cells;
				pp.recv_cells += t.ncells;
				pp.recv.push_back(to_dev(t, L.local_of[t.dst_box], -1));
			}
		}
		pp.d_send = pp.d_recv = nullptr;
		pp.send_off = send_cells_all;
		pp.recv_off = recv_cells_all;
		send_cells_all += pp.send_cells;
		recv_cells_all += pp.recv_cells;
		peers.push_back(pp);
	}
	bc.clear();
	for (size_t lb = 0; lb < L.local_ids.size(); ++lb)
		plan_bc(L, ng, (int)lb, L.boxes[L.local_ids[lb]], bc);
	max_tag_cells = 1;
	if (!need_device)
		return 0;
	std::vector<DevTag> dl;
	for (const HostTag &t : local) {
		dl.push_back(to_dev(t, L.local_of[t.dst_box], L.local_of[t.src_box]));
		max_tag_cells = std::max(max_tag_cells, t.ncells);
	}
	QK_CUDA((cudaError_t)upload(dl, &d_local));
	n_local = (int)dl.size();
	std::vector<DevTag> sa, ra;
	for (PeerPlan &pp : peers) {
		QK_CUDA((cudaError_t)upload(pp.send, &pp.d_send));
		QK_CUDA((cudaError_t)upload(pp.recv, &pp.d_recv));
		for (DevTag t : pp.send) {
			t.off += pp.send_off;
			max_send_tag = std::max(max_send_tag, t.ncells);
			sa.push_back(t);
		}
		for (DevTag t : pp.recv) {
			t.off += pp.recv_off;
			max_recv_tag = std::max(max_recv_tag, t.ncells);
			ra.push_back(t);
		}
	}
	QK_CUDA((cudaError_t)upload(sa, &d_send_all));
	QK_CUDA((cudaError_t)upload(ra, &d_recv_all));
	n_send_all = (int)sa.size();
	n_recv_all = (int)ra.size();
	std::vector<DevBcTag> db;
	for (const HostBcTag &t : bc) {
		DevBcTag d;
		d.box = t.box;
		for (int k = 0; k < 3; ++k) {
			d.lo[k] = t.region.lo[k];
			d.n[k] = blen(t.region, k);
		}
		db.push_back(d);
	}
	QK_CUDA((cudaError_t)upload(db, &d_bc));
	n_bc = (int)db.size();
	return 0;
}

void qk_exchange_plan::destroy()
{
	if (d_local)
		cudaFree(d_local);
	if (d_bc)
		cudaFree(d_bc);
	if (d_send_all)
		cudaFree(d_send_all);
	if (d_recv_all)
		cudaFree(d_recv_all);
	d_send_all = d_recv_all = nullptr;
	for (PeerPlan &pp : peers) {
		if (pp.d_send)
			cudaFree(pp.d_send);
		if (pp.d_recv)
			cudaFree(pp.d_recv);
	}
	d_local = nullptr;
	d_bc = nullptr;
	peers.clear();
}

extern "C" int qk_level_create(const qk_level_desc *desc, qk_level **out)
{
	if (!desc || !out || desc->nboxes_global <= 0 || desc->nghost < 0 || desc->ncomp <= 0)
		return QK_ERR_BAD_ARG;
	qk_level *L = new qk_level();
	L->domain = desc->domain;
	for (int d = 0; d < 3; ++d) {
		L->periodic[d] = desc->periodic[d];
		L->dx[d] = desc->dx[d];
	}
	L->nghost = desc->nghost;
	L->ncomp = desc->ncomp;
	L->my_rank = desc->my_rank;
	L->boxes.assign(desc->boxes_global, desc->boxes_global + desc->nboxes_global);
	L->owner.assign(desc->owner, desc->owner + desc->nboxes_global);
	L->bc_lo.assign(desc->bc_lo, desc->bc_lo + 3 * desc->ncomp);
	L->bc_hi.assign(desc->bc_hi, desc->bc_hi + 3 * desc->ncomp);
	L->local_of.assign(desc->nboxes_global, -1);
	L->nranks = 1;
	for (int b = 0; b < desc->nboxes_global; ++b) {
		L->nranks = std::max(L->nranks, L->owner[b] + 1);
		if (L->owner[b] == L->my_rank) {
			L->local_of[b] = (int)L->local_ids.size();
			L->local_ids.push_back(b);
			L->valid.push_back(L->boxes[b]);
		}
	}
	L->has_device = (qk_device_count() > 0);
	int rc = L->plan.build(*L, L->nghost, L->has_device);
	if (rc == 0)
		rc = L->plan1.build(*L, 1, L->has_device);
	if (rc == 0 && L->has_device) {
		rc = (int)upload(L->bc_lo, &L->d_bc_lo);
		if (rc == 0)
			rc = (int)upload(L->bc_hi, &L->d_bc_hi);
	}
	if (rc != 0) {
		qk_level_destroy(L);
		return rc;
	}
	*out = L;
	return 0;
}

extern "C" void qk_level_destroy(qk_level *L)
{
	if (!L)
		return;
	L->plan.destroy();
	L->plan1.destroy();
	qk_fused_free(L);
	qk_rad_free(L);
	L->free_scratch();
	if (L->d_bc_lo)
		cudaFree(L->d_bc_lo);
	if (L->d_bc_hi)
		cudaFree(L->d_bc_hi);
	L->ring.destroy();
	if (L->d_counters)
		cudaFree(L->d_counters);
	if (L->h_counters)
		cudaFreeHost(L->h_counters);
	if (L->send_all.p)
		cudaFree(L->send_all.p);
	if (L->recv_all.p)
		cudaFree(L->recv_all.p);
	if (L->comm_stream)
		cudaStreamDestroy(L->comm_stream);
	if (L->ev_packed)
		cudaEventDestroy(L->ev_packed);
	if (L->ev_received)
		cudaEventDestroy(L->ev_received);
	for (auto &kv : L->msg_send)
		cudaFree(kv.second.p);
	for (auto &kv : L->msg_recv)
		cudaFree(kv.second.p);
	delete L;
}

extern "C" int qk_level_nlocal(const qk_level *L) { return L ? (int)L->local_ids.size() : 0; }
extern "C" int qk_level_local_ids(const qk_level *L, int32_t *ids)
{
	if (!L || !ids)
		return QK_ERR_BAD_ARG;
	for (size_t i = 0; i < L->local_ids.size(); ++i)
		ids[i] = L->local_ids[i];
	return 0;
}
static int export_tags(const std::vector<HostTag> &src, qk_copy_tag *tags, int max_tags)
{
	const int n = (int)src.size();
	if (tags) {
		for (int i = 0; i < n && i < max_tags; ++i) {
			const HostTag &t = src[i];
			qk_copy_tag &o = tags[i];
			o.src_box = t.src_box;
			o.dst_box = t.dst_box;
			o.src_rank = t.src_rank;
			o.dst_rank = t.dst_rank;
			for (int d = 0; d < 3; ++d) {
				o.src_region.lo[d] = t.dst_region.lo[d] - t.shift[d];
				o.src_region.hi[d] = t.dst_region.hi[d] - t.shift[d];
				o.shift[d] = t.shift[d];
			}
			o.offset = t.offset;
			o.ncells = t.ncells;
		}
	}
	return n;
}
extern "C" int qk_level_remote_tags(const qk_level *L, qk_copy_tag *tags, int max_tags)
{
	return L ? export_tags(L->plan.remote, tags, max_tags) : QK_ERR_BAD_ARG;
}
extern "C" int qk_level_local_tags(const qk_level *L, qk_copy_tag *tags, int max_tags)
{
	return L ? export_tags(L->plan.local, tags, max_tags) : QK_ERR_BAD_ARG;
}

// ---- descriptor ring: per-call Array4 tables travel host -> device through pinned slots ----------------
int DescRing::init(size_t bytes_per_slot)
{
	if (h)
		return 0;
	slot_bytes = (bytes_per_slot + 255) & ~(size_t)255;
	QK_CUDA(cudaMallocHost(&h, slot_bytes * NSLOT));
	QK_CUDA(cudaMalloc(&d, slot_bytes * NSLOT));
	for (int i = 0; i < NSLOT; ++i) {
		QK_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
		used[i] = false;
	}
	return 0;
}
void DescRing::destroy()
{
	if (!h)
		return;
	for (int i = 0; i < NSLOT; ++i)
		cudaEventDestroy(ev[i]);
	cudaFreeHost(h);
	cudaFree(d);
	h = nullptr;
	d = nullptr;
}
void *DescRing::push(const void *src, size_t bytes, cudaStream_t s, int *err)
{
	*err = 0;
	if (bytes > slot_bytes) {
		*err = QK_ERR_BAD_ARG;
		return nullptr;
	}
	const int i = next;
	next = (next + 1) % NSLOT;
	if (used[i]) {
		cudaError_t e = cudaEventSynchronize(ev[i]);
		if (e != cudaSuccess) {
			*err = (int)e;
			return nullptr;
		}
	}
	memcpy(h + slot_bytes * i, src, bytes);
	cudaError_t e = cudaMemcpyAsync(d + slot_bytes * i, h + slot_bytes * i, bytes, cudaMemcpyHostToDevice, s);
	if (e == cudaSuccess)
		e = cudaEventRecord(ev[i], s);
	if (e != cudaSuccess) {
		*err = (int)e;
		return nullptr;
	}
	used[i] = true;
	return d + slot_bytes * i;
}

const A4 *qk_level::dev_table(const qk_array4 *arrs, cudaStream_t s, int *err)
{
	const int nb = (int)local_ids.size();
	*err = ring.init(sizeof(A4) * std::max(nb, 1) * 4);
	if (*err)
		return nullptr;
	std::vector<A4> t(nb);
	for (int b = 0; b < nb; ++b)
		t[b] = A4(arrs[b]);
	return (const A4 *)ring.push(t.data(), sizeof(A4) * nb, s, err);
}
const IA4 *qk_level::dev_table_int(const qk_iarray4 *arrs, cudaStream_t s, int *err)
{
	const int nb = (int)local_ids.size();
	*err = ring.init(sizeof(A4) * std::max(nb, 1) * 4);
	if (*err)
		return nullptr;
	std::vector<IA4> t(nb);
	for (int b = 0; b < nb; ++b)
		t[b] = IA4(arrs[b]);
	return (const IA4 *)ring.push(t.data(), sizeof(IA4) * nb, s, err);
}

static inline unsigned tag_blocks(int64_t max_cells) { return (unsigned)std::min<int64_t>(64, (max_cells + 255) / 256); }

#define QK_NEED_DEV(L)                                                                                                                               \
	do {                                                                                                                                         \
		if (!(L))                                                                                                                            \
			return QK_ERR_BAD_ARG;                                                                                                       \
		if (!(L)->has_device)                                                                                                                \
			return QK_ERR_NO_DEVICE;                                                                                                     \
	} while (0)

int qk_level::fill_local(const qk_exchange_plan &P, const A4 *tab, int scomp, int nc, cudaStream_t s)
{
	if (P.n_local == 0)
		return 0;
	dim3 grid(tag_blocks(P.max_tag_cells), P.n_local);
	k_tags<0><<<grid, 256, 0, s>>>(P.d_local, tab, scomp, nc, nullptr);
	QK_KERNEL_CHECK();
	return 0;
}
int qk_level::fill_bc(const qk_exchange_plan &P, const A4 *tab, int scomp, int nc, cudaStream_t s)
{
	if (P.n_bc == 0)
		return 0;
	dim3 grid(64, P.n_bc);
	k_phys_bc<<<grid, 256, 0, s>>>(P.d_bc, tab, scomp, nc, Box3(domain), d_bc_lo, d_bc_hi, periodic[0], periodic[1], periodic[2]);
	QK_KERNEL_CHECK();
	return 0;
}

extern "C" int qk_fill_boundary_local(qk_level *L, const qk_array4 *state, int scomp, int ncomp, void *stream)
{
	QK_NEED_DEV(L);
	int err;
	const A4 *tab = L->dev_table(state, S(stream), &err);
	if (err)
		return err;
	return L->fill_local(L->plan, tab, scomp, ncomp, S(stream));
}

static const PeerPlan *find_peer(const qk_exchange_plan &P, int peer)
{
	for (const PeerPlan &pp : P.peers)
		if (pp.peer == peer)
			return &pp;
	return nullptr;
}

extern "C" int qk_pack_ghosts(qk_level *L, int peer, const qk_array4 *state, int scomp, int ncomp, double *buf, int64_t *ndoubles, void *stream)
{
	QK_NEED_DEV(L);
	const PeerPlan *pp = find_peer(L->plan, peer);
	if (ndoubles)
		*ndoubles = pp ? pp->send_cells * ncomp : 0;
	if (!pp || pp->send.empty() || !buf)
		return 0;
	int err;
	const A4 *tab = L->dev_table(state, S(stream), &err);
	if (err)
		return err;
	int64_t mx = 1;
	for (const DevTag &t : pp->send)
		mx = std::max(mx, t.ncells);
	dim3 grid(tag_blocks(mx), (unsigned)pp->send.size());
	k_tags<1><<<grid, 256, 0, S(stream)>>>(pp->d_send, tab, scomp, ncomp, buf);
	QK_KERNEL_CHECK();
	return 0;
}

extern "C" int qk_unpack_ghosts(qk_level *L, int peer, const qk_array4 *state, int scomp, int ncomp, const double *buf, void *stream)
{
	QK_NEED_DEV(L);
	const PeerPlan *pp = find_peer(L->plan, peer);
	if (!pp || pp->recv.empty())
		return 0;
	int err;
	const A4 *tab = L->dev_table(state, S(stream), &err);
	if (err)
		return err;
	int64_t mx = 1;
	for (const DevTag &t : pp->recv)
		mx = std::max(mx, t.ncells);
	dim3 grid(tag_blocks(mx), (unsigned)pp->recv.size());
	k_tags<2><<<grid, 256, 0, S(stream)>>>(pp->d_recv, tab, scomp, ncomp, const_cast<double *>(buf));
	QK_KERNEL_CHECK();
	return 0;
}

extern "C" int qk_fill_physical_bc(qk_level *L, const qk_array4 *state, int scomp, int ncomp, void *stream)
{
	QK_NEED_DEV(L);
	int err;
	const A4 *tab = L->dev_table(state, S(stream), &err);
	if (err)
		return err;
	return L->fill_bc(L->plan, tab, scomp, ncomp, S(stream));
}

// fillBoundaryConditions on level 0 (src/simulation.hpp:1752-1765): same-rank copies, remote exchange over
// the attached communicator (NCCL grouped send/recv replacing MPI_Isend/Irecv), then the physical BCs.
extern "C" int qk_fill_boundary(qk_level *L, const qk_array4 *state, int scomp, int ncomp, void *stream)
{
	QK_NEED_DEV(L);
	int err;
	const A4 *tab = L->dev_table(state, S(stream), &err);
	if (err)
		return err;
	return L->fill_boundary_tab(tab, scomp, ncomp, S(stream));
}

int qk_level::fill_boundary_tab(const A4 *tab, int scomp, int nc, cudaStream_t s)
{
	ProfScope prof_("fill_boundary", s);
	// pack every peer's message with ONE launch, hand the grouped NCCL send/recv to the communication stream, run the same-rank
	// copies on the caller's stream meanwhile (NVLink traffic overlaps them), then unpack everything with ONE launch
	const bool remote = !plan.peers.empty();
	if (remote) {
		if (!comm)
			return QK_ERR_BAD_ARG; // multi-rank level without qk_level_set_comm
		const size_t ns = (size_t)plan.send_cells_all * nc * sizeof(double), nr = (size_t)plan.recv_cells_all * nc * sizeof(double);
		if (send_all.bytes < ns) {
			if (send_all.p)
				cudaFree(send_all.p);
			QK_CUDA(cudaMalloc(&send_all.p, ns));
			send_all.bytes = ns;
		}
		if (recv_all.bytes < nr) {
			if (recv_all.p)
				cudaFree(recv_all.p);
			QK_CUDA(cudaMalloc(&recv_all.p, nr));
			recv_all.bytes = nr;
		}
		if (!comm_stream) {
			QK_CUDA(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
			QK_CUDA(cudaEventCreateWithFlags(&ev_packed, cudaEventDisableTiming));
			QK_CUDA(cudaEventCreateWithFlags(&ev_received, cudaEventDisableTiming));
		}
		if (plan.n_send_all > 0) {
			dim3 grid(tag_blocks(plan.max_send_tag), (unsigned)plan.n_send_all);
			k_tags<1><<<grid, 256, 0, s>>>(plan.d_send_all, tab, scomp, nc, (double *)send_all.p);
			QK_KERNEL_CHECK();
		}
		QK_CUDA(cudaEventRecord(ev_packed, s));
		QK_CUDA(cudaStreamWaitEvent(comm_stream, ev_packed, 0));
		int rc = qk_comm_group_start(comm);
		if (rc)
			return rc;
		for (const PeerPlan &pp : plan.peers) {
			if (pp.recv_cells)
				rc = rc ? rc : qk_comm_recv(comm, (double *)recv_all.p + pp.recv_off * nc, (size_t)pp.recv_cells * nc * sizeof(double), pp.peer, comm_stream);
			if (pp.send_cells)
				rc = rc ? rc : qk_comm_send(comm, (double *)send_all.p + pp.send_off * nc, (size_t)pp.send_cells * nc * sizeof(double), pp.peer, comm_stream);
		}
		const int rc2 = qk_comm_group_end(comm);
		if (rc || rc2)
			return rc ? rc : rc2;
		QK_CUDA(cudaEventRecord(ev_received, comm_stream));
	}
	int rc = fill_local(plan, tab, scomp, nc, s);
	if (rc)
		return rc;
	if (remote) {
		QK_CUDA(cudaStreamWaitEvent(s, ev_received, 0));
		if (plan.n_recv_all > 0) {
			dim3 grid(tag_blocks(plan.max_recv_tag), (unsigned)plan.n_recv_all);
			k_tags<2><<<grid, 256, 0, s>>>(plan.d_recv_all, tab, scomp, nc, (double *)recv_all.p);
			QK_KERNEL_CHECK();
		}
	}
	return fill_bc(plan, tab, scomp, nc, s);
}

// redoFlag.FillBoundary(periodicity) (QuokkaSimulation.hpp:1157): 1-cell ghost layer of the int flags
int qk_level::fill_redo_flags(cudaStream_t s)
{
	int err;
	const IA4 *tab = dev_table_int(scr.redo.data(), s, &err);
	if (err)
		return err;
	const bool remote = !plan1.peers.empty();
	if (remote) {
		if (!comm)
			return QK_ERR_BAD_ARG;
		for (const PeerPlan &pp : plan1.peers) {
			MsgBuf &sb = msg_send[pp.peer], &rb = msg_recv[pp.peer];
			const size_t ns = (size_t)pp.send_cells * 4, nr = (size_t)pp.recv_cells * 4;
			if (sb.bytes < ns) {
				if (sb.p)
					cudaFree(sb.p);
				QK_CUDA(cudaMalloc(&sb.p, ns));
				sb.bytes = ns;
			}
			if (rb.bytes < nr) {
				if (rb.p)
					cudaFree(rb.p);
				QK_CUDA(cudaMalloc(&rb.p, nr));
				rb.bytes = nr;
			}
			if (!pp.send.empty()) {
				dim3 grid(8, (unsigned)pp.send.size());
				k_tags_int<1><<<grid, 256, 0, s>>>(pp.d_send, tab, (int32_t *)sb.p);
				QK_KERNEL_CHECK();
			}
		}
		int rc = qk_comm_group_start(comm);
		if (rc)
			return rc;
		for (const PeerPlan &pp : plan1.peers) {
			if (pp.recv_cells)
				rc = rc ? rc : qk_comm_recv(comm, msg_recv[pp.peer].p, (size_t)pp.recv_cells * 4, pp.peer, s);
			if (pp.send_cells)
				rc = rc ? rc : qk_comm_send(comm, msg_send[pp.peer].p, (size_t)pp.send_cells * 4, pp.peer, s);
		}
		const int rc2 = qk_comm_group_end(comm);
		if (rc || rc2)
			return rc ? rc : rc2;
	}
	if (plan1.n_local) {
		dim3 grid(8, plan1.n_local);
		k_tags_int<0><<<grid, 256, 0, s>>>(plan1.d_local, tab, nullptr);
		QK_KERNEL_CHECK();
	}
	if (remote) {
		for (const PeerPlan &pp : plan1.peers) {
			if (pp.recv.empty())
				continue;
			dim3 grid(8, (unsigned)pp.recv.size());
			k_tags_int<2><<<grid, 256, 0, s>>>(pp.d_recv, tab, (int32_t *)msg_recv[pp.peer].p);
			QK_KERNEL_CHECK();
		}
	}
	return 0;
}

extern "C" int qk_level_set_comm(qk_level *L, qk_comm *comm)
{
	if (!L)
		return QK_ERR_BAD_ARG;
	L->comm = comm;
	return 0;
}

// ------------------------------------------------------------------------------------------------
// scratch
// ------------------------------------------------------------------------------------------------
static qk_array4 mk_desc(double *p, const qk_box &b, int ncomp)
{
	qk_array4 a;
	a.p = p;
	a.jstride = blen(b, 0);
	a.kstride = a.jstride * blen(b, 1);
	a.nstride = a.kstride * blen(b, 2);
	for (int d = 0; d < 3; ++d) {
		a.begin[d] = b.lo[d];
		a.end[d] = b.hi[d] + 1;
	}
	a.ncomp = ncomp;
	return a;
}
static qk_box face_of(qk_box b, int d)
{
	b.hi[d] += 1;
	return b;
}

// pad_x: round the x pitch up to an even number of doubles so that every row starts 16-byte aligned (TMA bulk copies)
// ---- tensor-map descriptors (TMA tiles of the hydro and radiation sweeps, qk_march.cuh / qk_rad_kernels.cuh) ----------------------------------------------------------
typedef CUresult (*qk_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
				       const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static qk_encode_tiled_fn encode_tiled()
{
	static qk_encode_tiled_fn fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void *p = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
			fn = reinterpret_cast<qk_encode_tiled_fn>(p);
	}
	return fn;
}

// descriptor of a box {bx, by, bz, bc} of the FP64 array `a` viewed as a 4-D tensor (x, y, z, component); false when the array cannot be
// described (odd pitches, unaligned base): the caller then runs the kernels without TMA staging
bool qk_encode_tile(void *map128, const qk_array4 &a, unsigned bx, unsigned by, unsigned bz, unsigned bc)
{
	CUtensorMap *m = static_cast<CUtensorMap *>(map128);
	qk_encode_tiled_fn enc = encode_tiled();
	if (!enc || ((uintptr_t)a.p % 16) != 0 || (a.jstride % 2) != 0 || (a.kstride % 2) != 0 || (a.nstride % 2) != 0)
		return false;
	const cuuint64_t dims[4] = {(cuuint64_t)(a.end[0] - a.begin[0]), (cuuint64_t)(a.end[1] - a.begin[1]), (cuuint64_t)(a.end[2] - a.begin[2]),
				    (cuuint64_t)a.ncomp};
	const cuuint64_t strides[3] = {(cuuint64_t)a.jstride * 8, (cuuint64_t)a.kstride * 8, (cuuint64_t)a.nstride * 8};
	const cuuint32_t box[4] = {bx, by, bz, bc};
	const cuuint32_t es[4] = {1, 1, 1, 1};
	if (bc > dims[3])
		return false;
	return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, a.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
		   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


int qk_level::alloc_fabs(std::vector<qk_array4> &out, int ncomp, int grow, int face_dir, bool pad_x)
{
	const int nb = (int)valid.size();
	out.resize(nb);
	size_t total = 0;
	std::vector<size_t> off(nb);
	auto pitch = [&](const qk_box &bx) { return pad_x ? (size_t)((blen(bx, 0) + 1) & ~1) : (size_t)blen(bx, 0); };
	for (int b = 0; b < nb; ++b) {
		qk_box bx = bgrow(face_dir >= 0 ? face_of(valid[b], face_dir) : valid[b], grow);
		off[b] = total;
		size_t n = pitch(bx) * blen(bx, 1) * blen(bx, 2) * ncomp;
		total += (n + 31) & ~(size_t)31; // keep every FAB 256-byte aligned
	}
	total += 64; // slack: a bulk row copy may run a few elements past the last row
	double *p = nullptr;
	if (cudaMalloc(&p, total * sizeof(double)) != cudaSuccess) {
		cudaGetLastError();
		return QK_ERR_NOMEM;
	}
	scratch_ptrs.push_back(p);
	scratch_bytes += (int64_t)(total * sizeof(double));
	for (int b = 0; b < nb; ++b) {
		qk_box bx = bgrow(face_dir >= 0 ? face_of(valid[b], face_dir) : valid[b], grow);
		out[b] = mk_desc(p + off[b], bx, ncomp);
		if (pad_x) {
			out[b].jstride = (int64_t)pitch(bx);
			out[b].kstride = out[b].jstride * blen(bx, 1);
			out[b].nstride = out[b].kstride * blen(bx, 2);
		}
	}
	return 0;
}

int qk_level::ensure_counters()
{
	if (d_counters)
		return 0;
	QK_CUDA(cudaMalloc(&d_counters, 64 * sizeof(unsigned long long)));
	QK_CUDA(cudaMemset(d_counters, 0, 64 * sizeof(unsigned long long)));
	QK_CUDA(cudaMallocHost(&h_counters, 64 * sizeof(unsigned long long)));
	return 0;
}

int qk_level::ensure_faithful_scratch(int nv)
{
	if (scr.nv == nv)
		return 0;
	if (scr.nv != 0)
		return QK_ERR_UNSUPPORTED; // nvars changed between calls on one level
	const int nb = (int)valid.size();
	int rc = 0;
	rc = rc ? rc : alloc_fabs(scr.prim, nv, nghost, -1);
	rc = rc ? rc : alloc_fabs(scr.rhs, nv, 0, -1);
	for (int d = 0; d < 3 && !rc; ++d) {
		rc = rc ? rc : alloc_fabs(scr.chi[d], 1, 2, -1);
		rc = rc ? rc : alloc_fabs(scr.flx[d], nv, 0, d);
		rc = rc ? rc : alloc_fabs(scr.fvl[d], 1, 0, d);
		rc = rc ? rc : alloc_fabs(scr.frk[d], nv, 0, d);
		rc = rc ? rc : alloc_fabs(scr.avg[d], 1, 0, d);
		scr.faces[d].resize(nb);
		for (int b = 0; b < nb; ++b)
			scr.faces[d][b] = face_of(valid[b], d);
	}
	if (rc)
		return rc;
	// redoFlag: int, 1 ghost cell (QuokkaSimulation.hpp:1111)
	scr.redo.resize(nb);
	size_t total = 0;
	std::vector<size_t> off(nb);
	for (int b = 0; b < nb; ++b) {
		qk_box g = bgrow(valid[b], 1);
		off[b] = total;
		total += (((size_t)blen(g, 0) * blen(g, 1) * blen(g, 2)) + 63) & ~(size_t)63;
	}
	int32_t *p = nullptr;
	if (cudaMalloc(&p, total * 4) != cudaSuccess) {
		cudaGetLastError();
		return QK_ERR_NOMEM;
	}
	scratch_ptrs.push_back(p);
	scratch_bytes += (int64_t)total * 4;
	scr.redo_base = p;
	scr.redo_count = total;
	for (int b = 0; b < nb; ++b) {
		qk_box g = bgrow(valid[b], 1);
		qk_iarray4 a;
		a.p = p + off[b];
		a.jstride = blen(g, 0);
		a.kstride = a.jstride * blen(g, 1);
		a.nstride = a.kstride * blen(g, 2);
		for (int d = 0; d < 3; ++d) {
			a.begin[d] = g.lo[d];
			a.end[d] = g.hi[d] + 1;
		}
		a.ncomp = 1;
		scr.redo[b] = a;
	}
	scr.nv = nv;
	return 0;
}

int qk_level::ensure_fo_scratch()
{
	if (!scr.fof[0].empty())
		return 0;
	int rc = 0;
	for (int d = 0; d < 3 && !rc; ++d) {
		rc = rc ? rc : alloc_fabs(scr.fof[d], scr.nv, 0, d);
		rc = rc ? rc : alloc_fabs(scr.fov[d], 1, 0, d);
	}
	return rc;
}

void qk_level::free_scratch()
{
	for (void *p : scratch_ptrs)
		cudaFree(p);
	scratch_ptrs.clear();
	scratch_bytes = 0;
	scr = FaithfulScratch();
}

extern "C" int64_t qk_level_scratch_bytes(const qk_level *L) { return L ? L->scratch_bytes : 0; }

// ------------------------------------------------------------------------------------------------
// faithful stage
// ------------------------------------------------------------------------------------------------
#define QK_TRY(x)                                                                                                                                    \
	do {                                                                                                                                         \
		int r_ = (x);                                                                                                                        \
		if (r_ != 0)                                                                                                                         \
			return r_;                                                                                                                   \
	} while (0)

// sum over ranks of a host count: ParallelDescriptor::ReduceLongSum equivalent of redoFlag.sum (MultiFab::sum
// reduces over all ranks, QuokkaSimulation.hpp:1146)
int qk_level::global_sum(int64_t *v, cudaStream_t s)
{
	if (!comm || nranks == 1)
		return 0;
	return qk_comm_allreduce_sum_i64(comm, v, s);
}

// rhs + PdV + PredictStep on all local boxes from face arrays F,V; returns the GLOBAL redoFlag.sum()
int qk_level::update_from_fluxes(const qk_hydro_params *prm, std::vector<qk_array4> *F, std::vector<qk_array4> *V, const qk_array4 *U0,
				 const qk_array4 *Uout, double dt, int64_t *nbad, cudaStream_t s)
{
	const int nb = (int)valid.size();
	QK_TRY(qk_hydro_rhs_from_fluxes(nb, valid.data(), scr.rhs.data(), F[0].data(), F[1].data(), F[2].data(), dx, scr.nv, s));
	QK_TRY(qk_hydro_add_internal_energy_pdv(prm, nb, valid.data(), scr.rhs.data(), U0, dx, V[0].data(), V[1].data(), V[2].data(), scr.redo.data(), s));
	int64_t bad = 0;
	QK_TRY(qk_hydro_predict_step(prm, nb, valid.data(), U0, Uout, scr.rhs.data(), dt, scr.nv, scr.redo.data(), &bad, s));
	QK_TRY(global_sum(&bad, s));
	*nbad = bad;
	return 0;
}

// FOFC: computeFOHydroFluxes(U0) (QuokkaSimulation.hpp:1519-1557; evaluated lazily -- it depends on U0 only, so
// the values are those the reference computes up front at :1096), redoFlag.FillBoundary, replaceFluxes, redo.
int qk_level::fofc_redo(const qk_hydro_params *prm, std::vector<qk_array4> *F, std::vector<qk_array4> *V, const qk_array4 *U0,
			const qk_array4 *Uout, double dt, int64_t *nbad, cudaStream_t s)
{
	const int nb = (int)valid.size();
	QK_TRY(ensure_fo_scratch());
	if (!scr.fo_valid) {
		QK_TRY(qk_hydro_conserved_to_primitive(prm, nb, valid.data(), U0, scr.prim.data(), nghost, s));
		for (int d = 0; d < 3; ++d)
			QK_TRY(qk_hydro_flux_function(prm, 1, d, nb, valid.data(), scr.prim.data(), nullptr, nullptr, nullptr, scr.fof[d].data(), scr.fov[d].data(), s));
		scr.fo_valid = true;
	}
	QK_TRY(fill_redo_flags(s));
	for (int d = 0; d < 3; ++d) {
		QK_TRY(qk_hydro_replace_fluxes(d, nb, valid.data(), F[d].data(), scr.fof[d].data(), scr.redo.data(), scr.nv, s));
		QK_TRY(qk_hydro_replace_fluxes(d, nb, valid.data(), V[d].data(), scr.fov[d].data(), scr.redo.data(), 1, s));
	}
	return update_from_fluxes(prm, F, V, U0, Uout, dt, nbad, s);
}

// computeHydroFluxes (QuokkaSimulation.hpp:1403-1490) of state U into flx/fvl, then flux_rk2 += 0.5 F, avgFaceVel += 0.5 v (:1105-1108)
int qk_level::faithful_fluxes(const qk_hydro_params *prm, const qk_array4 *U, bool zero_rk, cudaStream_t s)
{
	const int nb = (int)valid.size();
	const int nv = scr.nv;
	QK_TRY(qk_hydro_conserved_to_primitive(prm, nb, valid.data(), U, scr.prim.data(), nghost, s));
	for (int d = 0; d < 3; ++d)
		QK_TRY(qk_hydro_flattening_coefficients(prm, d, nb, valid.data(), scr.prim.data(), scr.chi[d].data(), 2, s));
	for (int d = 0; d < 3; ++d)
		QK_TRY(qk_hydro_flux_function(prm, 0, d, nb, valid.data(), scr.prim.data(), scr.chi[0].data(), scr.chi[1].data(), scr.chi[2].data(),
					      scr.flx[d].data(), scr.fvl[d].data(), s));
	for (int d = 0; d < 3; ++d) {
		if (zero_rk) { // flux_rk2.setVal(0), avgFaceVel.setVal(0) (:1060-1073)
			QK_CUDA(cudaMemsetAsync(scr.frk[d][0].p, 0, (size_t)((char *)(scr.frk[d][nb - 1].p + scr.frk[d][nb - 1].nstride * nv) - (char *)scr.frk[d][0].p), s));
			QK_CUDA(cudaMemsetAsync(scr.avg[d][0].p, 0, (size_t)((char *)(scr.avg[d][nb - 1].p + scr.avg[d][nb - 1].nstride) - (char *)scr.avg[d][0].p), s));
		}
		QK_TRY(qk_saxpy(nb, scr.faces[d].data(), scr.frk[d].data(), 0.5, scr.flx[d].data(), nv, s));
		QK_TRY(qk_saxpy(nb, scr.faces[d].data(), scr.avg[d].data(), 0.5, scr.fvl[d].data(), 1, s));
	}
	return 0;
}

int qk_level::faithful_stage(const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage, const qk_array4 *Uout, double dt,
			     int64_t *ncells_bad, cudaStream_t s)
{
	const int nb = (int)valid.size();
	const int nv = 6 + prm->nscalars;
	QK_TRY(ensure_faithful_scratch(nv));
	if (stage == 1) {
		scr.fo_valid = false;
		QK_TRY(faithful_fluxes(prm, Ustage, true, s));
		scr.rk_valid = true;
	} else {
		if (!scr.rk_valid) { // stage 1 of this step ran on the fused path: rebuild 0.5 F(U0) first
			scr.fo_valid = false;
			QK_TRY(faithful_fluxes(prm, U0, true, s));
		}
		QK_TRY(faithful_fluxes(prm, Ustage, false, s));
		scr.rk_valid = false; // consumed
	}
	std::vector<qk_array4> *F = (stage == 1) ? scr.flx : scr.frk;
	std::vector<qk_array4> *V = (stage == 1) ? scr.fvl : scr.avg;
	QK_CUDA(cudaMemsetAsync(scr.redo_base, 0, scr.redo_count * 4, s)); // redoFlag.setVal(quokka::redoFlag::none)
	int64_t bad = 0;
	QK_TRY(update_from_fluxes(prm, F, V, U0, Uout, dt, &bad, s));
	scr.first_check_bad = bad;
	if (bad > 0)
		QK_TRY(fofc_redo(prm, F, V, U0, Uout, dt, &bad, s));
	if (bad == 0 || !prm->abort_on_fofc_failure) {
		QK_TRY(qk_hydro_enforce_limits(prm, nb, valid.data(), Uout, s));
		if (prm->use_dual_energy)
			QK_TRY(qk_hydro_sync_dual_energy(prm, nb, valid.data(), Uout, nullptr, s));
	}
	if (ncells_bad)
		*ncells_bad = bad;
	return 0;
}

extern "C" int qk_hydro_advance_stage_faithful(qk_level *L, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage,
					       const qk_array4 *Uout, double dt, int64_t *ncells_bad, void *stream)
{
	QK_NEED_DEV(L);
	if (!prm || (stage != 1 && stage != 2) || !U0 || !Ustage || !Uout)
		return QK_ERR_BAD_ARG;
	return L->faithful_stage(prm, stage, U0, Ustage, Uout, dt, ncells_bad, S(stream));
}

const std::vector<qk_array4> *qk_fused_kept_fluxes(const qk_level *L, int dir);

extern "C" int qk_level_stage_fluxes(const qk_level *L, int dir, qk_array4 *out)
{
	if (!L || !out || dir < 0 || dir > 2)
		return QK_ERR_BAD_ARG;
	if (const std::vector<qk_array4> *kept = qk_fused_kept_fluxes(L, dir)) { // the last stage ran on the fused flux-keeping path
		for (size_t b = 0; b < L->valid.size(); ++b)
			out[b] = (*kept)[b];
		return 0;
	}
	if (L->scr.nv == 0)
		return QK_ERR_BAD_ARG;
	for (size_t b = 0; b < L->valid.size(); ++b)
		out[b] = L->scr.flx[dir][b];
	return 0;
}

// The production entry point.  Until a level has a fused plan (qk_sweep.cu) it is the faithful path.
int qk_fused_stage(qk_level *L, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage, const qk_array4 *Uout, double dt,
		   int64_t *ncells_bad, cudaStream_t s, bool *handled, bool keepf);

static int advance_stage(qk_level *L, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage, const qk_array4 *Uout, double dt,
			 int64_t *ncells_bad, void *stream, bool keepf);

extern "C" int qk_hydro_advance_stage(qk_level *L, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage,
				      const qk_array4 *Uout, double dt, int64_t *ncells_bad, void *stream)
{
	return advance_stage(L, prm, stage, U0, Ustage, Uout, dt, ncells_bad, stream, false);
}

// The stage for a level with flux registers: the fused sweeps also store the stage's face fluxes (qk_sweep_keepf.cu); a stage the fused
// kernels cannot take, or one that flags a cell, runs on the faithful path, whose flux arrays qk_level_stage_fluxes then returns.
extern "C" int qk_hydro_advance_stage_keep_fluxes(qk_level *L, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage,
						  const qk_array4 *Uout, double dt, int64_t *ncells_bad, void *stream)
{
	return advance_stage(L, prm, stage, U0, Ustage, Uout, dt, ncells_bad, stream, true);
}

static int advance_stage(qk_level *L, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage, const qk_array4 *Uout, double dt,
			 int64_t *ncells_bad, void *stream, bool keepf)
{
	QK_NEED_DEV(L);
	if (!prm || (stage != 1 && stage != 2) || !U0 || !Ustage || !Uout)
		return QK_ERR_BAD_ARG;
	bool handled = false;
	QK_TRY(qk_fused_stage(L, prm, stage, U0, Ustage, Uout, dt, ncells_bad, S(stream), &handled, keepf));
	if (handled)
		return 0;
	int64_t bad = 0;
	if (stage == 1)
		L->scr.rk_valid = false;
	QK_TRY(L->faithful_stage(prm, stage, U0, Ustage, Uout, dt, &bad, S(stream)));
	if (L->scr.first_check_bad == 0)
		qk_fused_untaint(L); // every cell of Uout has rho > 0 again: the fused kernels may take the next stage
	if (ncells_bad)
		*ncells_bad = bad;
	return 0;
}
