// TEST INFRASTRUCTURE ONLY — the one piece of the reference that compiles stand-alone, used as a checker.
//
// This driver is ours; what it drives is the REFERENCE's own header, compiled where it lies:
//   /root/reference/frame_main/libs/MVS/Interface.h  (MVS::Interface, MVS::ARCHIVE::SerializeSave / SerializeLoad,
//   Interface::Platform::GetFullK / GetPose, MVS::HeaderDepthDataRaw)
// built by oracle/Makefile into oracle/_ref/mvsi_ref_tool (git-ignored; never linked into the product). tests/test_mvsi_ref.py
// uses it to pin the product's MVSI reader / writer and raw "DR" .dmap writer / reader (hc-mvs_b200/host/mvsi.cpp, densify.cpp)
// against files written / parsed by the reference's code. Nothing of the reference is copied into this repository.
//
// Exchange with the tests is a flat little-endian dump ("flat"): u32 counts, u32-length strings — see Load/SaveFlat below.
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <vector>
#include <limits>
#include <algorithm>
#include <string>
#include "Interface.h" // the reference's (include path set by oracle/Makefile)

using MVS::Interface;

namespace {
struct In {
	FILE* f; bool ok = true;
	template<typename T> void raw(T* p, size_t n) { if (ok && n) ok = fread(p, sizeof(T), n, f) == n; }
	uint32_t u32() { uint32_t v = 0; raw(&v, 1); return v; }
	std::string str() { std::string s(u32(), '\0'); if (!s.empty()) raw(&s[0], s.size()); return s; }
};
struct Out {
	FILE* f;
	template<typename T> void raw(const T* p, size_t n) { if (n) fwrite(p, sizeof(T), n, f); }
	void u32(uint32_t v) { raw(&v, 1); }
	void str(const std::string& s) { u32((uint32_t)s.size()); raw(s.data(), s.size()); }
};

bool LoadFlat(const char* file, Interface& obj) {
	FILE* f = fopen(file, "rb"); if (!f) return false;
	In in{f};
	obj.platforms.resize(in.u32());
	for (auto& p: obj.platforms) {
		p.name = in.str();
		p.cameras.resize(in.u32());
		for (auto& c: p.cameras) { c.name = in.str(); c.bandName = in.str(); c.width = in.u32(); c.height = in.u32(); in.raw(c.K.val, 9); in.raw(c.R.val, 9); in.raw(&c.C.x, 3); }
		p.poses.resize(in.u32());
		for (auto& q: p.poses) { in.raw(q.R.val, 9); in.raw(&q.C.x, 3); }
	}
	obj.images.resize(in.u32());
	for (auto& im: obj.images) { im.name = in.str(); im.maskName = in.str(); im.platformID = in.u32(); im.cameraID = in.u32(); im.poseID = in.u32(); im.ID = in.u32(); }
	obj.vertices.resize(in.u32());
	for (auto& v: obj.vertices) { in.raw(&v.X.x, 3); v.views.resize(in.u32()); for (auto& w: v.views) { w.imageID = in.u32(); in.raw(&w.confidence, 1); } }
	obj.verticesNormal.resize(in.u32()); for (auto& n: obj.verticesNormal) in.raw(&n.n.x, 3);
	obj.verticesColor.resize(in.u32()); for (auto& c: obj.verticesColor) in.raw(&c.c.x, 3);
	obj.lines.resize(in.u32());
	for (auto& l: obj.lines) { in.raw(&l.pt1.x, 3); in.raw(&l.pt2.x, 3); l.views.resize(in.u32()); for (auto& w: l.views) { w.imageID = in.u32(); in.raw(&w.confidence, 1); } }
	obj.linesNormal.resize(in.u32()); for (auto& n: obj.linesNormal) in.raw(&n.n.x, 3);
	obj.linesColor.resize(in.u32()); for (auto& c: obj.linesColor) in.raw(&c.c.x, 3);
	in.raw(obj.transform.val, 16);
	fclose(f);
	return in.ok;
}

bool SaveFlat(const char* file, const Interface& obj) {
	FILE* f = fopen(file, "wb"); if (!f) return false;
	Out out{f};
	out.u32((uint32_t)obj.platforms.size());
	for (const auto& p: obj.platforms) {
		out.str(p.name);
		out.u32((uint32_t)p.cameras.size());
		for (const auto& c: p.cameras) { out.str(c.name); out.str(c.bandName); out.u32(c.width); out.u32(c.height); out.raw(c.K.val, 9); out.raw(c.R.val, 9); out.raw(&c.C.x, 3); }
		out.u32((uint32_t)p.poses.size());
		for (const auto& q: p.poses) { out.raw(q.R.val, 9); out.raw(&q.C.x, 3); }
	}
	out.u32((uint32_t)obj.images.size());
	for (const auto& im: obj.images) { out.str(im.name); out.str(im.maskName); out.u32(im.platformID); out.u32(im.cameraID); out.u32(im.poseID); out.u32(im.ID); }
	out.u32((uint32_t)obj.vertices.size());
	for (const auto& v: obj.vertices) { out.raw(&v.X.x, 3); out.u32((uint32_t)v.views.size()); for (const auto& w: v.views) { out.u32(w.imageID); out.raw(&w.confidence, 1); } }
	out.u32((uint32_t)obj.verticesNormal.size()); for (const auto& n: obj.verticesNormal) out.raw(&n.n.x, 3);
	out.u32((uint32_t)obj.verticesColor.size()); for (const auto& c: obj.verticesColor) out.raw(&c.c.x, 3);
	out.u32((uint32_t)obj.lines.size());
	for (const auto& l: obj.lines) { out.raw(&l.pt1.x, 3); out.raw(&l.pt2.x, 3); out.u32((uint32_t)l.views.size()); for (const auto& w: l.views) { out.u32(w.imageID); out.raw(&w.confidence, 1); } }
	out.u32((uint32_t)obj.linesNormal.size()); for (const auto& n: obj.linesNormal) out.raw(&n.n.x, 3);
	out.u32((uint32_t)obj.linesColor.size()); for (const auto& c: obj.linesColor) out.raw(&c.c.x, 3);
	out.raw(obj.transform.val, 16);
	const bool ok = ferror(f) == 0;
	fclose(f);
	return ok;
}
} // namespace

int main(int argc, char** argv) {
	const std::string cmd = argc > 1 ? argv[1] : "";
	if (cmd == "from-flat" && argc == 5) { // flat -> .mvs through the reference's SerializeSave
		Interface obj;
		if (!LoadFlat(argv[2], obj)) return 2;
		return MVS::ARCHIVE::SerializeSave(obj, argv[3], (uint32_t)atoi(argv[4])) ? 0 : 3;
	}
	if (cmd == "to-flat" && argc == 4) { // .mvs -> flat through the reference's SerializeLoad; prints the stream version
		Interface obj; uint32_t version = 0;
		if (!MVS::ARCHIVE::SerializeLoad(obj, argv[2], &version)) return 3;
		printf("%u\n", version);
		return SaveFlat(argv[3], obj) ? 0 : 2;
	}
	if (cmd == "cams" && argc == 4) {
		// per calibrated image with a stated resolution: Interface::Platform::GetFullK and GetPose (Interface.h:437-458, 608-617)
		Interface obj;
		if (!MVS::ARCHIVE::SerializeLoad(obj, argv[2])) return 3;
		FILE* f = fopen(argv[3], "wb"); if (!f) return 2;
		for (uint32_t i=0; i<(uint32_t)obj.images.size(); ++i) {
			const Interface::Image& im = obj.images[i];
			if (!im.IsValid()) continue;
			const Interface::Platform& pl = obj.platforms[im.platformID];
			const Interface::Platform::Camera& cam = pl.cameras[im.cameraID];
			if (!cam.HasResolution()) continue;
			const Interface::Mat33d K = pl.GetFullK(im.cameraID, cam.width, cam.height);
			const Interface::Platform::Pose pose = obj.GetPose(i);
			fwrite(&i, 4, 1, f); fwrite(K.val, 8, 9, f); fwrite(pose.R.val, 8, 9, f); fwrite(&pose.C.x, 8, 3, f);
		}
		fclose(f);
		return 0;
	}
	if (cmd == "dmap-layout" && argc == 2) {
		// the raw depth-data header as the reference declares it (Interface.h:634-652)
		MVS::HeaderDepthDataRaw h;
		printf("%zu %u %d %d %d %zu %zu %zu %zu %zu %zu %zu\n", sizeof(h), (unsigned)MVS::HeaderDepthDataRaw::HeaderDepthDataRawName(),
			(int)MVS::HeaderDepthDataRaw::HAS_DEPTH, (int)MVS::HeaderDepthDataRaw::HAS_NORMAL, (int)MVS::HeaderDepthDataRaw::HAS_CONF,
			offsetof(MVS::HeaderDepthDataRaw, type), offsetof(MVS::HeaderDepthDataRaw, imageWidth), offsetof(MVS::HeaderDepthDataRaw, imageHeight),
			offsetof(MVS::HeaderDepthDataRaw, depthWidth), offsetof(MVS::HeaderDepthDataRaw, depthHeight),
			offsetof(MVS::HeaderDepthDataRaw, dMin), offsetof(MVS::HeaderDepthDataRaw, dMax));
		return 0;
	}
	if (cmd == "dmap-header" && argc == 3) { // read a file's header INTO the reference's struct and print its fields
		FILE* f = fopen(argv[2], "rb"); if (!f) return 2;
		MVS::HeaderDepthDataRaw h;
		const bool ok = fread(&h, sizeof(h), 1, f) == 1;
		fclose(f);
		if (!ok) return 3;
		printf("%d %u %u %u %u %u %.9g %.9g\n", h.name == MVS::HeaderDepthDataRaw::HeaderDepthDataRawName() ? 1 : 0, (unsigned)h.type,
			h.imageWidth, h.imageHeight, h.depthWidth, h.depthHeight, h.dMin, h.dMax);
		return 0;
	}
	if (cmd == "dmap-write-header" && argc == 10) { // file type w h dw dh dMin dMax: a header written FROM the reference's struct
		FILE* f = fopen(argv[2], "wb"); if (!f) return 2;
		MVS::HeaderDepthDataRaw h;
		h.name = MVS::HeaderDepthDataRaw::HeaderDepthDataRawName(); h.type = (uint8_t)atoi(argv[3]); h.padding = 0;
		h.imageWidth = (uint32_t)atoi(argv[4]); h.imageHeight = (uint32_t)atoi(argv[5]); h.depthWidth = (uint32_t)atoi(argv[6]); h.depthHeight = (uint32_t)atoi(argv[7]);
		h.dMin = (float)atof(argv[8]); h.dMax = (float)atof(argv[9]);
		fwrite(&h, sizeof(h), 1, f);
		fclose(f);
		return 0;
	}
	fprintf(stderr, "usage: mvsi_ref_tool from-flat <flat> <out.mvs> <version> | to-flat <in.mvs> <flat> | cams <in.mvs> <out.bin> | dmap-layout | dmap-header <file> | dmap-write-header <file> type w h dw dh dMin dMax\n");
	return 1;
}
