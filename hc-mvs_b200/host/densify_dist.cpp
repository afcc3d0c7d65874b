// Scene::DenseReconstruction over `world` GPUs of one box, one process (one hcmvs_ctx with a communicator) per GPU.
//
// The reference (SceneDensify.cpp:3532-3574) is single-node CPU code; what it offers is that EstimateDepthMap of one view only
// reads static inputs and writes that view's maps, so the scene shards by reference view:
//   * every rank selects the neighbour views of 1/world of the images (Scene::SelectNeighborViews, the reference's
//     `#pragma omp parallel for` over images, :3652-3667) and the results are all-gathered — small host records;
//   * every rank uploads 1/world of the images over PCIe and receives the others over NVLink (HCMVS_EXCHANGE_IMAGES);
//   * views are dealt round-robin in FuseDepthMaps' connection order (:3286-3303); round s of the estimation is broadcast in place
//     while round s+1 runs; the views of the last, incomplete round are estimated in row bands by ALL ranks
//     (hcmvs_estimate_depthmap_rows: bit-identical to the whole-view estimate);
//   * FilterDepthMap of a view runs on its owner, the pending results are exchanged and committed everywhere (:4146-4178);
//   * FuseDepthMaps is sequential over views by definition and runs on rank 0, which ends up with Scene::densecloud.
// Results are identical to the single-GPU call: the kernels are deterministic and the RNG is counter-based on (seed, view, pixel).
#include "densify.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <cstdio>
#include <cstdlib>

namespace hcmvs_host {

static double NowD() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

ShardPlan MakeShardPlan(const std::vector<uint32_t>& validViews, const std::vector<uint32_t>& nScoredNeighbors, int world, bool splitRows) {
	ShardPlan p; p.world = std::max(world, 1);
	p.order = validViews;
	std::stable_sort(p.order.begin(), p.order.end(), [&](uint32_t a, uint32_t b) {
		if (nScoredNeighbors[a] != nScoredNeighbors[b]) return nScoredNeighbors[a] > nScoredNeighbors[b];
		return a < b;
	});
	const size_t n = p.order.size();
	p.splitRows = splitRows && p.world > 1 && n%(size_t)p.world != 0;
	p.wholeRounds = (int)(p.splitRows ? n/(size_t)p.world : (n+(size_t)p.world-1)/(size_t)p.world);
	return p;
}
std::vector<uint32_t> ShardPlan::SplitViews() const {
	if (!splitRows) return {};
	return std::vector<uint32_t>(order.begin()+(size_t)wholeRounds*world, order.end());
}
std::vector<uint32_t> ShardPlan::WholeViewsOf(int rank) const {
	std::vector<uint32_t> out;
	const size_t n = std::min(order.size(), (size_t)wholeRounds*world);
	for (size_t k=0; k<n; ++k) if ((int)(k%(size_t)world) == rank) out.push_back(order[k]);
	return out;
}
std::vector<uint32_t> ShardPlan::ViewsOf(int rank) const { // the views a rank filters (round-robin over the WHOLE order, split views included)
	std::vector<uint32_t> out;
	for (size_t k=0; k<order.size(); ++k) if ((int)(k%(size_t)world) == rank) out.push_back(order[k]);
	return out;
}
std::vector<int32_t> ShardPlan::RoundOwners(int round, size_t nViews) const {
	std::vector<int32_t> o(nViews, -1);
	for (size_t k=(size_t)round*world; k<std::min(order.size(), (size_t)(round+1)*world); ++k) o[order[k]] = (int32_t)(k%(size_t)world);
	return o;
}
std::vector<int32_t> ShardPlan::SplitOwners(size_t nViews) const {
	std::vector<int32_t> o(nViews, -1);
	for (uint32_t v: SplitViews()) o[v] = HCMVS_OWNER_SPLIT_ROWS;
	return o;
}
std::vector<int32_t> ShardPlan::Owners(size_t nViews, const std::vector<char>* only) const {
	std::vector<int32_t> o(nViews, -1);
	for (size_t k=0; k<order.size(); ++k) if (!only || (*only)[order[k]]) o[order[k]] = (int32_t)(k%(size_t)world);
	return o;
}
void ShardPlan::RowsOf(int rank, int height, int& r0, int& r1) const { r0 = (int)((long long)rank*height/world); r1 = (int)((long long)(rank+1)*height/world); }


struct DistributedReconstruction::Impl {
	Scene& scene; hcmvs_ctx* ctx; hcmvs_params P; ViewSelectionParams VS; int rank, world;
	DepthMapsData data;
	ShardPlan plan; std::vector<uint32_t> valid, mineWhole, split, mine;
	struct Init { std::vector<float> depth; std::vector<double> vertices; std::vector<uint32_t> tris; };
	std::vector<Init> init;     // initial maps (host) of the views this rank estimates
	std::vector<uint32_t> est;  // those views in estimation order: this rank's whole views by round, then the row-split views
	std::vector<char> ok;       // per image: SelectViews succeeded (own + split views after SelectOne, all views after Gather)
	std::vector<char> inited;   // per image: InitViews done on this rank
	bool lazy = false;          // one-call job: selection + initial maps run on worker threads AHEAD of the estimation inside Run()
	bool gathered = false;
	DenseReconstructionStats st;
	std::string err;
	// lazy jobs: the selection + initial maps of the FIRST view this rank estimates (what its GPU waits for) start with Prepare(), on
	// their own threads, so that they overlap the image uploads and the NVLink image exchange instead of following them
	std::thread firstSel; int firstSelRes = 0;
	Impl(Scene& s, hcmvs_ctx* c, const hcmvs_params& p, const ViewSelectionParams& vs, int r, int w): scene(s), ctx(c), P(p), VS(vs), rank(r), world(w), data(s, c, p, vs) {}
	~Impl() { if (firstSel.joinable()) firstSel.join(); }
	bool fail(const std::string& m) { err = "rank "+std::to_string(rank)+": "+m; return false; }
	bool lib(const char* what) { return fail(std::string(what)+": "+hcmvs_last_error()); }
	int Owner(uint32_t i) const { return (int)(i%(uint32_t)world); } // uploads the image, selects its neighbours, estimates it (unless row-split), filters it
	unsigned Threads(size_t work) const { const unsigned hc = std::thread::hardware_concurrency(); return (unsigned)std::max<size_t>(1, std::min<size_t>((hc > 1 ? hc-1 : 1u)/(unsigned)world+1u, work)); }
	bool Prepare();
	bool SelectOne(uint32_t i, unsigned threads = 1);   // SelectViews + the initial maps of a view this rank estimates (worker threads)
	bool Gather();                // all-gather of the selection results; neighbour lists of every view on this rank
	bool InitOne(uint32_t i);     // InitViews (neighbour lists to the device)
	bool UploadOne(uint32_t i);
	bool UploadInitial();
	bool Run(uint64_t seed, bool runFilter, bool download);
};

bool DistributedReconstruction::Impl::SelectOne(uint32_t i, unsigned threads) {
	ok[i] = data.SelectViews(i, threads) ? 1 : 0;
	if (!ok[i]) return true;
	DepthData& dd = data.arrDepthData[i];
	if (P.nMinViewsTrustPoint >= 2) { // SceneDensify.cpp:781-812
		if (!TriangulateInit(scene, i, dd.points, true, init[i].vertices, init[i].tris, dd.dMin, dd.dMax)) return false;
		dd.dMin *= 0.9f; dd.dMax *= 1.1f;
	} else SparseInitDepth(scene, i, dd.points, init[i].depth, dd.dMin, dd.dMax);
	return true;
}

bool DistributedReconstruction::Impl::InitOne(uint32_t i) {
	if (inited[i] || !ok[i]) return true;
	inited[i] = 1;
	if (!data.InitViews(i, P.nNumViews)) { if (!data.lastError.empty()) return fail(data.lastError); data.arrDepthData[i].valid = false; ok[i] = 0; }
	return true;
}

bool DistributedReconstruction::Impl::UploadOne(uint32_t i) {
	DepthData& dd = data.arrDepthData[i];
	if (P.nMinViewsTrustPoint >= 2) {
		if (hcmvs_init_depthmap_triangles(ctx, i, init[i].vertices.data(), (int)(init[i].vertices.size()/3), init[i].tris.data(), (int)(init[i].tris.size()/3), dd.dMin, dd.dMax) != HCMVS_OK) return lib("hcmvs_init_depthmap_triangles");
		st.h2dBytes += (uint64_t)init[i].vertices.size()*8+(uint64_t)init[i].tris.size()*4;
	} else {
		if (hcmvs_init_depthmap(ctx, i, init[i].depth.data(), nullptr, dd.dMin, dd.dMax) != HCMVS_OK) return lib("hcmvs_init_depthmap");
		st.h2dBytes += (uint64_t)init[i].depth.size()*4;
	}
	return true;
}

bool DistributedReconstruction::Impl::UploadInitial() {
	for (uint32_t i: est) if (ok[i] && !UploadOne(i)) return false;
	return true;
}

bool DistributedReconstruction::Impl::Prepare() {
	st = DenseReconstructionStats();
	const uint32_t nImages = (uint32_t)scene.images.size();
	const double t0 = NowD();
	for (Image& im: scene.images) { im.camera.ComposeP(); im.neighbors.clear(); }
	if (hcmvs_begin_scene(ctx) != HCMVS_OK) return lib("hcmvs_begin_scene"); // initial-map uploads overlap the running estimation (see the header)
	for (DepthData& dd: data.arrDepthData) dd = DepthData();
	ok.assign(nImages, 0); inited.assign(nImages, 0); init.assign(nImages, Init()); gathered = false;
	// The plan needs nothing but the image count: views in INDEX order, view i whole on rank i % world in round i / world, the last
	// nImages % world views in row bands on every rank. (FuseDepthMaps' connection order, :3286-3303, only matters to the fusion itself;
	// dealing by index lets a rank start estimating its first view while the other views are still being selected.)
	{
		std::vector<uint32_t> all(nImages), zeros(nImages, 0);
		for (uint32_t i=0; i<nImages; ++i) all[i] = i;
		plan = MakeShardPlan(all, zeros, world, true);
	}
	mineWhole = plan.WholeViewsOf(rank); split = plan.SplitViews(); mine = plan.ViewsOf(rank);
	est = mineWhole; est.insert(est.end(), split.begin(), split.end());
	if (firstSel.joinable()) firstSel.join(); // a Prepare() whose Run() never came
	firstSelRes = 0;
	if (lazy && !est.empty()) firstSel = std::thread([this]() { firstSelRes = SelectOne(est[0], Threads(64)+1) ? 1 : -1; });
	// colour travels with the gray image; a rank only knows it for the images it holds, and the scene is homogeneous in that respect
	uint32_t hasColor = 0;
	for (uint32_t i=0; i<nImages; ++i) if (Owner(i) == rank && !scene.images[i].bgr.empty()) hasColor = 1;
	{
		std::vector<uint32_t> all((size_t)world);
		if (hcmvs_comm_allgather_host(ctx, &hasColor, all.data(), sizeof(uint32_t)) != HCMVS_OK) return lib("hcmvs_comm_allgather_host");
		hasColor = *std::max_element(all.begin(), all.end());
	}
	// ---- images: this rank's share over PCIe, everything to every rank over NVLink behind it (the broadcasts wait for each upload's event)
	std::atomic<size_t> next{0}; std::atomic<bool> bad{false};
	std::vector<std::thread> pool;
	if (!lazy) // staged job: select + make the initial maps of this rank's views now, while this thread moves the images
		for (unsigned t=0; t<Threads(est.size()); ++t) pool.emplace_back([&]() { size_t k; while ((k = next.fetch_add(1)) < est.size()) if (!SelectOne(est[k])) bad.store(true); });
	bool upOk = true; std::string upErr;
	for (uint32_t i=0; i<nImages && upOk; ++i) {
		Image& im = scene.images[i];
		if (Owner(i) == rank) {
			if (im.gray.empty() && im.bgr.empty()) { upOk = false; upErr = "image "+std::to_string(i)+" has no pixels on the rank that uploads it"; break; }
			if (!data.UploadView(i)) { upOk = false; upErr = data.lastError; break; }
			st.h2dBytes += (uint64_t)im.width*im.height*(4+(im.bgr.empty() ? 0 : 3));
		} else {
			if (hcmvs_set_view_remote(ctx, i, im.width, im.height, im.camera.K, im.camera.R, im.camera.C, (int)hasColor) != HCMVS_OK) { upOk = false; upErr = std::string("hcmvs_set_view_remote: ")+hcmvs_last_error(); break; }
			data.arrDepthData[i].uploaded = true;
		}
	}
	if (upOk) {
		std::vector<int32_t> owner(nImages);
		for (uint32_t i=0; i<nImages; ++i) owner[i] = Owner(i);
		if (hcmvs_exchange_maps(ctx, owner.data(), nImages, HCMVS_EXCHANGE_IMAGES) != HCMVS_OK) { upOk = false; upErr = std::string("hcmvs_exchange_maps(images): ")+hcmvs_last_error(); }
	}
	for (std::thread& th: pool) th.join();
	if (!upOk) return fail(upErr);
	if (bad.load()) return fail("cannot triangulate the sparse points of a view");
	if (!lazy) {
		for (uint32_t i: est) if (!InitOne(i)) return false; // before the gather: a view whose InitViews fails is reported as not valid
		if (!Gather()) return false;
		st.secSelect = NowD()-t0;
	}
	if (getenv("HCMVS_DIST_DEBUG")) fprintf(stderr, "[dist %d] prepare (%s) %.3f s\n", rank, lazy ? "images only; selection rides the estimation" : "images + selection + gather", NowD()-t0);
	return true;
}

namespace {
// what one rank tells the others about a view it selected
struct SelHeader { uint32_t valid, nScored, nNeighbors, nPoints; float avgDepth; uint32_t pad[3]; };
static_assert(sizeof(SelHeader) == 32, "record layout");
}

bool DistributedReconstruction::Impl::Gather() {
	// all-gather of the selection results (header + neighbour list; every rank needs them for the filter decisions, rank 0 for the fusion),
	// then the neighbour lists of every valid view go to the device
	if (gathered) return true;
	const uint32_t nImages = (uint32_t)scene.images.size();
	const size_t perView = sizeof(SelHeader)+HCMVS_MAX_FUSE_VIEWS*sizeof(ViewScore);
	const size_t slots = (nImages+(uint32_t)world-1)/(uint32_t)world; // views a rank owns at most
	std::vector<char> send(perView*slots, 0), recv(perView*slots*(size_t)world);
	for (uint32_t i=0; i<nImages; ++i) {
		if (Owner(i) != rank) continue;
		char* rec = send.data()+perView*(size_t)(i/(uint32_t)world);
		const DepthData& dd = data.arrDepthData[i];
		SelHeader h; std::memset(&h, 0, sizeof(h));
		h.valid = ok[i]; h.nScored = (uint32_t)scene.images[i].neighbors.size(); h.avgDepth = scene.images[i].avgDepth;
		h.nNeighbors = (uint32_t)std::min<size_t>(dd.neighbors.size(), HCMVS_MAX_FUSE_VIEWS);
		std::memcpy(rec, &h, sizeof(h));
		if (h.nNeighbors) std::memcpy(rec+sizeof(h), dd.neighbors.data(), h.nNeighbors*sizeof(ViewScore));
	}
	if (hcmvs_comm_allgather_host(ctx, send.data(), recv.data(), send.size()) != HCMVS_OK) return lib("hcmvs_comm_allgather_host");
	const std::vector<uint32_t> splitViews = plan.SplitViews();
	for (uint32_t i=0; i<nImages; ++i) {
		if (Owner(i) == rank) continue;
		if (std::find(splitViews.begin(), splitViews.end(), i) != splitViews.end()) continue; // every rank selected the row-split views itself
		const char* rec = recv.data()+send.size()*(size_t)Owner(i)+perView*(size_t)(i/(uint32_t)world);
		SelHeader h; std::memcpy(&h, rec, sizeof(h));
		DepthData& dd = data.arrDepthData[i];
		ok[i] = (char)h.valid; dd.valid = h.valid != 0;
		scene.images[i].avgDepth = h.avgDepth;
		scene.images[i].neighbors.assign(h.nScored, ViewScore()); // only its size is read from here on (FuseDepthMaps' connection order)
		dd.neighbors.resize(h.nNeighbors); if (h.nNeighbors) std::memcpy(dd.neighbors.data(), rec+sizeof(h), h.nNeighbors*sizeof(ViewScore));
	}
	valid.clear();
	for (uint32_t i=0; i<nImages; ++i) { if (!InitOne(i)) return false; if (ok[i]) valid.push_back(i); }
	if (valid.empty()) return fail("no image has enough neighbour views");
	gathered = true;
	return true;
}

bool DistributedReconstruction::Impl::Run(uint64_t seed, bool runFilter, bool download) {
	const uint32_t nImages = (uint32_t)scene.images.size();
	const double tRun = NowD();
	// one-call job: selection + initial maps are made by workers, in estimation order, while this thread feeds the GPU
	std::vector<std::atomic<int>> ready(lazy ? est.size() : 0);
	for (auto& r: ready) r.store(0);
	std::atomic<size_t> nextSel{0};
	std::vector<std::thread> pool;
	struct Joiner { std::vector<std::thread>& p; ~Joiner() { for (std::thread& th: p) if (th.joinable()) th.join(); } } joiner{pool};
	if (lazy) {
		// the other views this rank OWNS but does not estimate whole (none: owners estimate their views) need no work here; the row-split
		// views are selected by every rank
		// the first view is what this rank's GPU waits for: all of the rank's share of the cores work inside its selection first
		if (firstSel.joinable()) firstSel.join();
		if (!est.empty()) { ready[0].store(firstSelRes != 0 ? firstSelRes : (SelectOne(est[0], Threads(64)+1) ? 1 : -1), std::memory_order_release); nextSel.store(1); }
		for (unsigned t=0; t<Threads(est.size()); ++t) pool.emplace_back([&]() { size_t k; while ((k = nextSel.fetch_add(1)) < est.size()) ready[k].store(SelectOne(est[k]) ? 1 : -1, std::memory_order_release); });
	}
	size_t nextEst = 0; // position in `est`
	auto prepareView = [&](uint32_t i) -> bool { // neighbour lists + initial maps of a view about to be estimated; false + empty err: the view is not valid
		if (lazy) {
			int r; while ((r = ready[nextEst].load(std::memory_order_acquire)) == 0) std::this_thread::yield();
			if (r < 0) return fail("cannot triangulate the sparse points of a view");
			if (est[nextEst] != i) return fail("internal: estimation order");
			++nextEst;
			if (!InitOne(i)) return false;
			if (ok[i]) { if (!UploadOne(i)) return false; std::vector<float>().swap(init[i].depth); std::vector<double>().swap(init[i].vertices); std::vector<uint32_t>().swap(init[i].tris); }
		}
		return true;
	};
	const double t1 = NowD();
	// ---- estimation: round s = the s-th view of every rank, broadcast in place behind round s+1; then the row-split views
	for (unsigned it=0; it<std::max(1u, P.nEstimationIters_external); ++it) { // SceneDensify.cpp:3684
		if (it > 0 && P.viewspread && hcmvs_snapshot_maps(ctx) != HCMVS_OK) return lib("hcmvs_snapshot_maps");
		for (int s=0; s<plan.wholeRounds; ++s) {
			if ((size_t)s < mineWhole.size()) {
				const uint32_t i = mineWhole[(size_t)s];
				if (it == 0 && !prepareView(i)) return false;
				if (ok[i]) { if (hcmvs_estimate_depthmap(ctx, i, (int)it, seed) != HCMVS_OK) return lib("hcmvs_estimate_depthmap"); }
				else if (it == 0 && hcmvs_alloc_depthmap(ctx, i) != HCMVS_OK) return lib("hcmvs_alloc_depthmap"); // a view without neighbours: empty maps keep the rounds aligned
			}
			const std::vector<int32_t> own = plan.RoundOwners(s, nImages);
			if (hcmvs_exchange_maps(ctx, own.data(), nImages, HCMVS_EXCHANGE_ESTIMATED|HCMVS_EXCHANGE_ASYNC) != HCMVS_OK) return lib("hcmvs_exchange_maps");
		}
		for (uint32_t i: split) {
			if (it == 0 && !prepareView(i)) return false;
			int r0, r1; plan.RowsOf(rank, scene.images[i].height, r0, r1);
			if (ok[i]) { if (hcmvs_estimate_depthmap_rows(ctx, i, (int)it, seed, r0, r1) != HCMVS_OK) return lib("hcmvs_estimate_depthmap_rows"); }
			else if (it == 0 && hcmvs_alloc_depthmap(ctx, i) != HCMVS_OK) return lib("hcmvs_alloc_depthmap");
		}
		if (!split.empty()) {
			const std::vector<int32_t> own = plan.SplitOwners(nImages);
			if (hcmvs_exchange_maps(ctx, own.data(), nImages, HCMVS_EXCHANGE_ESTIMATED|HCMVS_EXCHANGE_ASYNC) != HCMVS_OK) return lib("hcmvs_exchange_maps");
		}
		if (hcmvs_exchange_wait(ctx) != HCMVS_OK) return lib("hcmvs_exchange_wait");
	}
	if (lazy) {
		for (std::thread& th: pool) if (th.joinable()) th.join();
		if (!Gather()) return false; // off the critical path: the GPU is still busy with the last rounds
		st.secSelect = NowD()-tRun;
	}
	if (hcmvs_sync(ctx) != HCMVS_OK) return lib("hcmvs_sync"); // attribution of the stage times only: the filter kernels queue behind the estimation anyway
	const double t3 = NowD(); st.secEstimate = t3-t1;
	// ---- FilterDepthMap on the owners (neighbours with a depth map, at most 8; SceneDensify.cpp:4093-4185), exchange, commit
	if (runFilter) {
		std::vector<char> filtered(nImages, 0);
		auto neighboursOf = [&](uint32_t i) {
			const DepthData& dd = data.arrDepthData[i];
			std::vector<uint32_t> idxNb;
			for (uint32_t k=0; k<dd.neighbors.size() && idxNb.size() < 8; ++k) if (ok[dd.neighbors[k].ID]) idxNb.push_back(k);
			return idxNb;
		};
		for (uint32_t i: valid) if (neighboursOf(i).size() >= std::min(P.nMinViewsFilter, scene.nCalibratedImages()-1)) filtered[i] = 1; // the same decision on every rank
		for (uint32_t i: mine) {
			if (!filtered[i]) continue;
			if (!data.FilterDepthMap(i, neighboursOf(i), P.bFilterAdjust != 0)) return fail(data.lastError);
		}
		const std::vector<int32_t> own = plan.Owners(nImages, &filtered);
		if (hcmvs_exchange_maps(ctx, own.data(), nImages, HCMVS_EXCHANGE_FILTERED) != HCMVS_OK) return lib("hcmvs_exchange_maps(filtered)");
		if (hcmvs_commit_filtered(ctx) != HCMVS_OK) return lib("hcmvs_commit_filtered");
	}
	const double t4 = NowD(); st.secFilter = t4-t3;
	// ---- FuseDepthMaps (sequential over views by definition) on rank 0
	scene.densecloud = PointCloud();
	if (rank == 0) {
		if (download) {
			if (!data.FuseDepthMaps(scene.densecloud, true, true)) return fail(data.lastError);
			st.d2hBytes += (uint64_t)scene.densecloud.size()*(12+12+3+4)+(uint64_t)scene.densecloud.views.size()*8;
		} else if (hcmvs_fuse_depthmaps(ctx, 1, 1, nullptr) != HCMVS_OK) return lib("hcmvs_fuse_depthmaps"); // the cloud stays in HBM
	}
	if (hcmvs_sync(ctx) != HCMVS_OK) return lib("hcmvs_sync");
	st.secFuse = NowD()-t4;
	if (getenv("HCMVS_DIST_DEBUG")) fprintf(stderr, "[dist %d] run: estimate %.3f filter %.3f fuse %.3f s\n", rank, st.secEstimate, st.secFilter, st.secFuse);
	return true;
}

DistributedReconstruction::DistributedReconstruction(Scene& scene, hcmvs_ctx* ctx, const hcmvs_params& P, const ViewSelectionParams& VS, int rank, int world)
	: impl(new Impl(scene, ctx, P, VS, rank, world)) {}
DistributedReconstruction::~DistributedReconstruction() { delete impl; }
bool DistributedReconstruction::Prepare(bool lazy) { impl->lazy = lazy; return impl->Prepare(); }
bool DistributedReconstruction::UploadInitial() { return impl->UploadInitial(); }
bool DistributedReconstruction::Run(uint64_t seed, bool runFilter, bool download) { return impl->Run(seed, runFilter, download); }
const std::string& DistributedReconstruction::Error() const { return impl->err; }
const DenseReconstructionStats& DistributedReconstruction::Stats() const { return impl->st; }
const ShardPlan& DistributedReconstruction::Plan() const { return impl->plan; }
size_t DistributedReconstruction::ValidViews() const { return impl->valid.size(); }

bool DenseReconstructionDistributed(Scene& scene, hcmvs_ctx* ctx, const hcmvs_params& P, const ViewSelectionParams& VS, uint64_t seed, bool runFilter,
	int rank, int world, DenseReconstructionStats* stats, std::string* err)
{
	if (world < 1 || rank < 0 || rank >= world) { if (err) *err = "bad rank / world"; return false; }
	if (world == 1) return DenseReconstruction(scene, ctx, P, VS, seed, runFilter, std::string(), stats, err);
	DistributedReconstruction job(scene, ctx, P, VS, rank, world);
	const bool ok = job.Prepare(true) && job.Run(seed, runFilter, true);
	if (!ok && err) *err = job.Error();
	if (stats) *stats = job.Stats();
	return ok;
}

} // namespace hcmvs_host
