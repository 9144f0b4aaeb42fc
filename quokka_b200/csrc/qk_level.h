// qk_level.h -- internal definition of the opaque qk_level (include/quokka_b200.h) and its helpers.
#pragma once
#include "qk_common.cuh"

#include <map>
#include <vector>

struct qk_comm; // qk_comm.cpp

// one box<->box ghost copy, in the DESTINATION box's index space (src index = dst index - shift)
struct HostTag {
	int32_t src_box, dst_box; // global box ids
	int32_t src_rank, dst_rank;
	qk_box dst_region;
	int32_t shift[3];
	int64_t ncells;
	int64_t offset; // cells before this tag in its (src_rank -> dst_rank) message
};
struct HostBcTag {
	int32_t box; // local index
	qk_box region;
};
struct DevTag {
	int32_t dst, src; // LOCAL box indices (or -1 on the remote side)
	int32_t lo[3], n[3], sh[3];
	int32_t pad;
	int64_t ncells, off;
};
struct DevBcTag {
	int32_t box;
	int32_t lo[3], n[3];
};
struct PeerPlan {
	int peer;
	std::vector<DevTag> send, recv;
	int64_t send_cells, recv_cells; // per component
	int64_t send_off, recv_off;	// cells (per component) before this peer's message in the level's combined send / recv buffers
	DevTag *d_send, *d_recv;
};

struct qk_level;
struct qk_exchange_plan {
	int nghost = 0;
	std::vector<HostTag> local, remote;
	std::vector<PeerPlan> peers;
	std::vector<HostBcTag> bc;
	DevTag *d_local = nullptr;
	DevBcTag *d_bc = nullptr;
	// all peers' pack / unpack tags in one table each (off = position in the combined buffer): one launch packs every message
	DevTag *d_send_all = nullptr, *d_recv_all = nullptr;
	int n_send_all = 0, n_recv_all = 0;
	int64_t send_cells_all = 0, recv_cells_all = 0, max_send_tag = 1, max_recv_tag = 1;
	int n_local = 0, n_bc = 0;
	int64_t max_tag_cells = 1;
	int build(const qk_level &L, int ng, bool need_device);
	void destroy();
};

struct DescRing {
	static const int NSLOT = 64;
	char *h = nullptr, *d = nullptr;
	size_t slot_bytes = 0;
	cudaEvent_t ev[NSLOT];
	bool used[NSLOT];
	int next = 0;
	int init(size_t bytes_per_slot);
	void destroy();
	void *push(const void *src, size_t bytes, cudaStream_t s, int *err);
};

struct MsgBuf {
	void *p = nullptr;
	size_t bytes = 0;
};

// scratch of the faithful (materialised-flux) stage path
struct FaithfulScratch {
	int nv = 0;
	std::vector<qk_array4> prim, rhs, chi[3], flx[3], fvl[3], frk[3], avg[3], fof[3], fov[3];
	std::vector<qk_box> faces[3];
	std::vector<qk_iarray4> redo;
	int32_t *redo_base = nullptr;
	size_t redo_count = 0;
	bool fo_valid = false;
	bool rk_valid = false;	      // frk/avg hold 0.5*F(U0), 0.5*faceVel(U0) of the CURRENT step (set by a faithful stage 1)
	int64_t first_check_bad = 0; // redoFlag.sum() before FOFC in the last faithful stage
};

struct FusedState; // qk_sweep.cu
struct RadState;   // qk_rad.cu

struct qk_level {
	qk_box domain;
	int periodic[3];
	double dx[3];
	int nghost, ncomp, my_rank, nranks;
	std::vector<qk_box> boxes; // global BoxArray
	std::vector<int32_t> owner, bc_lo, bc_hi;
	std::vector<int32_t> local_ids, local_of;
	std::vector<qk_box> valid; // local boxes in MFIter order
	bool has_device = false;
	qk_exchange_plan plan;	// nghost-wide state exchange
	qk_exchange_plan plan1; // 1-cell exchange (redoFlag)
	int32_t *d_bc_lo = nullptr, *d_bc_hi = nullptr;
	DescRing ring;
	qk_comm *comm = nullptr;
	std::map<int, MsgBuf> msg_send, msg_recv; // redoFlag exchange (rare)
	MsgBuf send_all, recv_all;		   // combined ghost messages of the state exchange
	cudaStream_t comm_stream = nullptr;	   // NCCL send/recv run here while the same-rank copies run on the caller's stream
	cudaEvent_t ev_packed = nullptr, ev_received = nullptr;
	unsigned long long *d_counters = nullptr, *h_counters = nullptr;

	std::vector<void *> scratch_ptrs;
	int64_t scratch_bytes = 0;
	FaithfulScratch scr;
	FusedState *fused = nullptr;
	RadState *rad = nullptr;

	const A4 *dev_table(const qk_array4 *arrs, cudaStream_t s, int *err);
	const IA4 *dev_table_int(const qk_iarray4 *arrs, cudaStream_t s, int *err);
	int fill_local(const qk_exchange_plan &P, const A4 *tab, int scomp, int nc, cudaStream_t s);
	int fill_bc(const qk_exchange_plan &P, const A4 *tab, int scomp, int nc, cudaStream_t s);
	int fill_boundary_tab(const A4 *tab, int scomp, int nc, cudaStream_t s);
	int fill_redo_flags(cudaStream_t s);
	int alloc_fabs(std::vector<qk_array4> &out, int ncomp, int grow, int face_dir, bool pad_x = false);
	int ensure_counters();
	int ensure_faithful_scratch(int nv);
	int ensure_fo_scratch();
	void free_scratch();
	int global_sum(int64_t *v, cudaStream_t s);
	int update_from_fluxes(const qk_hydro_params *prm, std::vector<qk_array4> *F, std::vector<qk_array4> *V, const qk_array4 *U0, const qk_array4 *Uout,
			       double dt, int64_t *nbad, cudaStream_t s);
	int fofc_redo(const qk_hydro_params *prm, std::vector<qk_array4> *F, std::vector<qk_array4> *V, const qk_array4 *U0, const qk_array4 *Uout, double dt,
		      int64_t *nbad, cudaStream_t s);
	int faithful_fluxes(const qk_hydro_params *prm, const qk_array4 *U, bool zero_rk, cudaStream_t s);
	int faithful_stage(const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage, const qk_array4 *Uout, double dt,
			   int64_t *ncells_bad, cudaStream_t s);
};

void qk_plan_tags(const qk_level &L, int ng, std::vector<HostTag> &out);
void qk_fused_free(qk_level *L);
void qk_rad_free(qk_level *L);
void qk_fused_untaint(qk_level *L);
int qk_fused_max_signal(qk_level *L, const qk_hydro_params *prm, const qk_array4 *state, double out[2], cudaStream_t s);

// ---- communicator (qk_comm.cpp): NCCL resolved at run time with dlopen, so the library loads on machines
// without NCCL and never conflicts with the copy a host application (or torch) already loaded ----
extern "C" {
int qk_comm_group_start(qk_comm *c);
int qk_comm_group_end(qk_comm *c);
int qk_comm_send(qk_comm *c, const void *buf, size_t bytes, int peer, cudaStream_t s);
int qk_comm_recv(qk_comm *c, void *buf, size_t bytes, int peer, cudaStream_t s);
int qk_comm_allreduce_sum_i64(qk_comm *c, int64_t *v, cudaStream_t s);
int qk_comm_allreduce_max_f64(qk_comm *c, double *v, cudaStream_t s);
// CUtensorMap (128 bytes at map128) of a box {bx, by, bz, bc} of the FP64 array `a` viewed as a 4-D tensor (x, y, z, component); false when the
// array cannot be described (odd pitches, unaligned base, no driver entry point): the caller then runs its kernels without TMA staging
bool qk_encode_tile(void *map128, const qk_array4 &a, unsigned bx, unsigned by, unsigned bz, unsigned bc);
int qk_comm_allreduce_dev_u64(qk_comm *c, unsigned long long *d_vals, int count, int is_max, cudaStream_t s);
}
