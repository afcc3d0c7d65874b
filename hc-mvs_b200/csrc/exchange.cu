// Multi-GPU map exchange inside the C ABI: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The scene shards by reference view: every rank estimates / filters its own views and the others need those
// maps for FilterDepthMap and FuseDepthMaps. Each view's buffers are broadcast IN PLACE from its owner — all broadcasts of
// one exchange in a single NCCL group, so they run as one fused transfer without staging copies — which makes the exchange
// usable from a plain C++ host (the reference's language) with no Python / torch in the process.
//
// NCCL is loaded lazily with dlopen: a single-GPU user never needs the library, and in a process that already loaded an NCCL
// (e.g. through torch) the same SONAME resolves to that copy.
#include "hcmvs_internal.h"
#include <dlfcn.h>
#include <cstring>
#include <vector>
#include <nccl.h>

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { hcmvs_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return HCMVS_ERR_CUDA; } } while (0)

namespace {
struct Nccl {
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr; // NCCL >= 2.18; optional
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
};
Nccl g_nccl;

int LoadNccl() {
	if (g_nccl.ok) return HCMVS_OK;
	const char* names[] = {"libnccl.so.2", "libnccl.so"};
	for (const char* n: names) { g_nccl.lib = dlopen(n, RTLD_NOW|RTLD_GLOBAL); if (g_nccl.lib) break; }
	if (!g_nccl.lib) { hcmvs_set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror()); return HCMVS_ERR_UNSUPPORTED; }
	#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name); if (!g_nccl.field) { hcmvs_set_error("NCCL symbol %s missing", name); return HCMVS_ERR_UNSUPPORTED; }
	SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
	SYM(Broadcast, "ncclBroadcast") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
	SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather")
	#undef SYM
	*(void**)(&g_nccl.CommSplit) = dlsym(g_nccl.lib, "ncclCommSplit");
	g_nccl.ok = true;
	return HCMVS_OK;
}
}

#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { hcmvs_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); return HCMVS_ERR_CUDA; } } while (0)

extern "C" int hcmvs_comm_unique_id(void* id128) {
	if (!id128) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	int r = LoadNccl(); if (r) return r;
	static_assert(HCMVS_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
	ncclUniqueId id;
	NK(g_nccl.GetUniqueId(&id));
	std::memcpy(id128, &id, sizeof(id));
	return HCMVS_OK;
}

extern "C" int hcmvs_comm_init(hcmvs_ctx* ctx, const void* id128, int rank, int world) {
	if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) { hcmvs_set_error("bad communicator arguments (rank %d of %d)", rank, world); return HCMVS_ERR_ARG; }
	int r = LoadNccl(); if (r) return r;
	cudaSetDevice(ctx->device);
	if (ctx->comm) { g_nccl.CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
	ncclUniqueId id; std::memcpy(&id, id128, sizeof(id));
	ncclComm_t comm;
	NK(g_nccl.CommInitRank(&comm, world, id, rank));
	ctx->comm = comm; ctx->rank = rank; ctx->world = world;
	// a second communicator + stream for the small agreement that precedes every exchange (status, depth ranges): it must not queue
	// behind the asynchronous map broadcasts of the data communicator, which wait for the compute stream
	ctx->ctlComm = nullptr;
	if (g_nccl.CommSplit) { ncclComm_t ctl; if (g_nccl.CommSplit(comm, 0, rank, &ctl, nullptr) == ncclSuccess) ctx->ctlComm = ctl; }
	if (!ctx->ctlStream) CK(cudaStreamCreateWithFlags(&ctx->ctlStream, cudaStreamNonBlocking));
	return HCMVS_OK;
}

void hcmvs_comm_release(hcmvs_ctx* ctx) {
	if (ctx->commStream) { cudaStreamSynchronize(ctx->commStream); cudaStreamDestroy(ctx->commStream); cudaEventDestroy(ctx->commDone); cudaEventDestroy(ctx->commReady); ctx->commStream = nullptr; }
	if (ctx->ctlStream) { cudaStreamSynchronize(ctx->ctlStream); cudaStreamDestroy(ctx->ctlStream); ctx->ctlStream = nullptr; }
	cudaFree(ctx->ctl_d); ctx->ctl_d = nullptr; ctx->ctlCap = 0;
	if (ctx->ctlComm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->ctlComm);
	if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
	ctx->comm = nullptr; ctx->ctlComm = nullptr;
}

extern "C" int hcmvs_exchange_wait(hcmvs_ctx* ctx) {
	if (!ctx) { hcmvs_set_error("null context"); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	if (ctx->commPending) {
		CK(cudaEventRecord(ctx->commDone, ctx->commStream));
		CK(cudaStreamWaitEvent(ctx->stream, ctx->commDone, 0)); // later compute work sees the received maps
		ctx->commPending = false;
	}
	return HCMVS_OK;
}

extern "C" int hcmvs_exchange_maps(hcmvs_ctx* ctx, const int32_t* owner, uint32_t n_views, int what) {
	if (!ctx || !owner) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	const bool async = (what & HCMVS_EXCHANGE_ASYNC) != 0;
	what &= ~HCMVS_EXCHANGE_ASYNC;
	if (what != HCMVS_EXCHANGE_ESTIMATED && what != HCMVS_EXCHANGE_FILTERED && what != HCMVS_EXCHANGE_IMAGES) { hcmvs_set_error("unknown exchange kind %d", what); return HCMVS_ERR_ARG; }
	if (!ctx->comm) { hcmvs_set_error("no communicator (call hcmvs_comm_init)"); return HCMVS_ERR_STATE; }
	if (what == HCMVS_EXCHANGE_ESTIMATED) for (View& v: ctx->views) v.depthValid = false; // received maps replace dn
	if (n_views > ctx->views.size()) { hcmvs_set_error("owner list longer than the scene (%u > %zu views)", n_views, ctx->views.size()); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	// Phase 1 — local checks and allocations. Nothing collective has been issued yet, so a failing rank must not simply return: its
	// peers would enter the broadcasts and wait for it forever. The outcome is agreed on first (phase 2).
	int status = HCMVS_OK;
	auto fail = [&](int code) { if (status == HCMVS_OK) status = code; };
	const size_t nCtl = (size_t)n_views*2+1;   // per view (dMin, dMax) contributed by its owner, then the number of failing ranks
	std::vector<float> ctl(nCtl, 0.f);
	for (uint32_t i=0; i<n_views && status == HCMVS_OK; ++i) {
		View& v = ctx->views[i];
		if (owner[i] == HCMVS_OWNER_SPLIT_ROWS) {
			// a view estimated in row bands by every rank (hcmvs_estimate_depthmap_rows): each rank contributes its band in place
			if (what != HCMVS_EXCHANGE_ESTIMATED) { hcmvs_set_error("row-split ownership only applies to the estimated maps (view %u)", i); fail(HCMVS_ERR_ARG); }
			else if (!v.set || !v.hasMaps || !v.dn_d || !v.conf_d) { hcmvs_set_error("rank %d holds no maps of the row-split view %u", ctx->rank, i); fail(HCMVS_ERR_STATE); }
			continue;
		}
		if (owner[i] < 0) continue;
		if (owner[i] >= ctx->world) { hcmvs_set_error("view %u is owned by rank %d of %d", i, owner[i], ctx->world); fail(HCMVS_ERR_ARG); continue; }
		if (!v.set) { hcmvs_set_error("view %u not set on rank %d (every rank holds every image)", i, ctx->rank); fail(HCMVS_ERR_STATE); continue; }
		const size_t n = (size_t)v.w*v.h;
		cudaError_t ce = cudaSuccess;
		if (what == HCMVS_EXCHANGE_IMAGES) {
			if (!v.img_d) { hcmvs_set_error("rank %d: view %u has no image buffers", ctx->rank, i); fail(HCMVS_ERR_STATE); }
			continue;
		}
		if (owner[i] == ctx->rank) {
			if (!v.hasMaps || !v.dn_d) { hcmvs_set_error("rank %d owns view %u but has no maps for it", ctx->rank, i); fail(HCMVS_ERR_STATE); }
			else if (what == HCMVS_EXCHANGE_FILTERED && !v.hasFiltered) { hcmvs_set_error("rank %d owns view %u but has not filtered it", ctx->rank, i); fail(HCMVS_ERR_STATE); }
			ctl[2*i] = v.dMin; ctl[2*i+1] = v.dMax;
		} else if (what == HCMVS_EXCHANGE_ESTIMATED) {
			if (!v.dn_d) ce = cudaMalloc(&v.dn_d, n*sizeof(float4));
			if (ce == cudaSuccess && !v.conf_d) ce = cudaMalloc(&v.conf_d, n*4);
		} else {
			if (!v.fdepth_d) ce = cudaMalloc(&v.fdepth_d, n*4);
			if (ce == cudaSuccess && !v.fconf_d) ce = cudaMalloc(&v.fconf_d, n*4);
		}
		if (ce != cudaSuccess) { hcmvs_set_error("rank %d: receive buffers of view %u: %s", ctx->rank, i, cudaGetErrorString(ce)); fail(HCMVS_ERR_CUDA); }
	}
	// Phase 2 — agreement: one small all-reduce (sum) on the control communicator / stream carries the failure count and, for the
	// estimated maps, every view's depth range from its owner (the receivers need it for .dmap output and later range checks).
	ctl[nCtl-1] = status == HCMVS_OK ? 0.f : 1.f;
	{
		ncclComm_t cc = (ncclComm_t)(ctx->ctlComm ? ctx->ctlComm : ctx->comm);
		cudaStream_t cs = ctx->ctlComm ? ctx->ctlStream : ctx->stream;
		if (ctx->ctlCap < nCtl) { if (ctx->ctl_d) { cudaStreamSynchronize(cs); cudaFree(ctx->ctl_d); ctx->ctl_d = nullptr; } CK(cudaMalloc(&ctx->ctl_d, nCtl*sizeof(float))); ctx->ctlCap = nCtl; }
		CK(cudaMemcpyAsync(ctx->ctl_d, ctl.data(), nCtl*sizeof(float), cudaMemcpyHostToDevice, cs));
		NK(g_nccl.AllReduce(ctx->ctl_d, ctx->ctl_d, nCtl, ncclFloat, ncclSum, cc, cs));
		CK(cudaMemcpyAsync(ctl.data(), ctx->ctl_d, nCtl*sizeof(float), cudaMemcpyDeviceToHost, cs));
		CK(cudaStreamSynchronize(cs));
	}
	if (status != HCMVS_OK) return status;
	if (ctl[nCtl-1] != 0.f) { hcmvs_set_error("map exchange abandoned: %d rank(s) failed their checks (this rank, %d, passed)", (int)ctl[nCtl-1], ctx->rank); return HCMVS_ERR_STATE; }
	// Phase 3 — the broadcasts: every rank walks the same list in the same order, so they pair up by position.
	// async: they run on the communication stream behind everything queued on the compute stream so far, and overlap
	// whatever is queued on the compute stream afterwards (the next view's sweeps) until hcmvs_exchange_wait
	cudaStream_t st = ctx->stream;
	if (async) {
		if (!ctx->commStream) { CK(cudaStreamCreateWithFlags(&ctx->commStream, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&ctx->commDone, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->commReady, cudaEventDisableTiming)); }
		CK(cudaEventRecord(ctx->commReady, ctx->stream));
		CK(cudaStreamWaitEvent(ctx->commStream, ctx->commReady, 0));
		st = ctx->commStream; ctx->commPending = true;
	} else hcmvs_time_begin(ctx, ST_EXCHANGE);
	// inside the group the first error is remembered and the group is ALWAYS closed: an open group would swallow every later NCCL call
	if (what == HCMVS_EXCHANGE_IMAGES) // the images this rank uploaded (copy stream) must have landed before they are sent
		for (uint32_t i=0; i<n_views; ++i) if (owner[i] == ctx->rank && ctx->views[i].imgReady) CK(cudaStreamWaitEvent(st, ctx->views[i].imgReady, 0));
	ncclResult_t first = ncclSuccess; const char* firstWhat = "";
	#define NG(call) do { if (first == ncclSuccess) { const ncclResult_t r_ = (call); if (r_ != ncclSuccess) { first = r_; firstWhat = #call; } } } while (0)
	NK(g_nccl.GroupStart());
	for (uint32_t i=0; i<n_views; ++i) {
		if (owner[i] == HCMVS_OWNER_SPLIT_ROWS) {
			View& v = ctx->views[i];
			for (int r=0; r<ctx->world; ++r) {
				const size_t row0 = (size_t)r*v.h/ctx->world, row1 = (size_t)(r+1)*v.h/ctx->world, cnt = (row1-row0)*v.w;
				if (!cnt) continue;
				NG(g_nccl.Broadcast(v.dn_d+row0*v.w, v.dn_d+row0*v.w, cnt*4, ncclFloat, r, (ncclComm_t)ctx->comm, st));
				NG(g_nccl.Broadcast(v.conf_d+row0*v.w, v.conf_d+row0*v.w, cnt, ncclFloat, r, (ncclComm_t)ctx->comm, st));
			}
			continue;
		}
		if (owner[i] < 0) continue;
		View& v = ctx->views[i];
		const size_t n = (size_t)v.w*v.h;
		if (what == HCMVS_EXCHANGE_IMAGES) {
			NG(g_nccl.Broadcast(v.img_d, v.img_d, n, ncclFloat, owner[i], (ncclComm_t)ctx->comm, st));
			if (v.bgr_d) NG(g_nccl.Broadcast(v.bgr_d, v.bgr_d, n*3, ncclUint8, owner[i], (ncclComm_t)ctx->comm, st));
		} else if (what == HCMVS_EXCHANGE_ESTIMATED) {
			NG(g_nccl.Broadcast(v.dn_d, v.dn_d, n*4, ncclFloat, owner[i], (ncclComm_t)ctx->comm, st));
			NG(g_nccl.Broadcast(v.conf_d, v.conf_d, n, ncclFloat, owner[i], (ncclComm_t)ctx->comm, st));
		} else {
			NG(g_nccl.Broadcast(v.fdepth_d, v.fdepth_d, n, ncclFloat, owner[i], (ncclComm_t)ctx->comm, st));
			NG(g_nccl.Broadcast(v.fconf_d, v.fconf_d, n, ncclFloat, owner[i], (ncclComm_t)ctx->comm, st));
		}
	}
	{ const ncclResult_t r_ = g_nccl.GroupEnd(); if (first == ncclSuccess && r_ != ncclSuccess) { first = r_; firstWhat = "ncclGroupEnd"; } }
	#undef NG
	if (!async) hcmvs_time_end(ctx);
	if (first != ncclSuccess) { hcmvs_set_error("%s -> %s", firstWhat, g_nccl.GetErrorString(first)); return HCMVS_ERR_CUDA; }
	if (what == HCMVS_EXCHANGE_IMAGES) {
		// the received gray images feed the gather textures (block-linear arrays); whoever samples them waits on imgReady
		for (uint32_t i=0; i<n_views; ++i) {
			if (owner[i] < 0 || owner[i] == ctx->rank) continue;
			View& v = ctx->views[i];
			CK(cudaMemcpy2DToArrayAsync(v.arr, 0, 0, v.img_d, (size_t)v.w*4, (size_t)v.w*4, v.h, cudaMemcpyDeviceToDevice, st));
			v.graValid = false;
			if (!v.imgReady) CK(cudaEventCreateWithFlags(&v.imgReady, cudaEventDisableTiming));
			CK(cudaEventRecord(v.imgReady, st));
		}
		return HCMVS_OK;
	}
	for (uint32_t i=0; i<n_views; ++i) {
		if (owner[i] < 0 || owner[i] == ctx->rank) continue;
		View& v = ctx->views[i];
		if (what == HCMVS_EXCHANGE_ESTIMATED) { v.hasMaps = true; v.dMin = ctl[2*i]; v.dMax = ctl[2*i+1]; } // the owner's depth range travels with the maps
		else v.hasFiltered = true; // hcmvs_commit_filtered applies it on this rank too
	}
	return HCMVS_OK;
}

extern "C" int hcmvs_comm_allgather_host(hcmvs_ctx* ctx, const void* send, void* recv, uint64_t bytes_per_rank) {
	if (!ctx || !send || !recv || !bytes_per_rank) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	if (!ctx->comm) { hcmvs_set_error("no communicator (call hcmvs_comm_init)"); return HCMVS_ERR_STATE; }
	cudaSetDevice(ctx->device);
	ncclComm_t cc = (ncclComm_t)(ctx->ctlComm ? ctx->ctlComm : ctx->comm);
	cudaStream_t cs = ctx->ctlComm ? ctx->ctlStream : ctx->stream;
	const size_t total = bytes_per_rank*(size_t)ctx->world;
	char* buf = nullptr;
	CK(cudaMalloc(&buf, total));
	cudaError_t e = cudaMemcpyAsync(buf+bytes_per_rank*(size_t)ctx->rank, send, bytes_per_rank, cudaMemcpyHostToDevice, cs);
	ncclResult_t nr = ncclSuccess;
	if (e == cudaSuccess) nr = g_nccl.AllGather(buf+bytes_per_rank*(size_t)ctx->rank, buf, bytes_per_rank, ncclUint8, cc, cs);
	if (e == cudaSuccess && nr == ncclSuccess) e = cudaMemcpyAsync(recv, buf, total, cudaMemcpyDeviceToHost, cs);
	if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
	cudaFree(buf);
	if (nr != ncclSuccess) { hcmvs_set_error("ncclAllGather -> %s", g_nccl.GetErrorString(nr)); return HCMVS_ERR_CUDA; }
	if (e != cudaSuccess) { hcmvs_set_error("hcmvs_comm_allgather_host: %s", cudaGetErrorString(e)); return HCMVS_ERR_CUDA; }
	return HCMVS_OK;
}
