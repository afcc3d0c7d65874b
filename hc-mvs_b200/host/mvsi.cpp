// MVSI project reader / writer, camera composition and the image decoders — see mvsi.h.
#include "mvsi.h"
#include "densify.h"
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <zlib.h>

namespace hcmvs_host {

// ------------------------------------------------------------------------------------------------ archive primitives
namespace {
struct Reader {
	FILE* f; bool ok = true;
	template<typename T> void raw(T* p, size_t n) { if (ok && n) ok = fread(p, sizeof(T), n, f) == n; }
	uint32_t u32() { uint32_t v = 0; raw(&v, 1); return v; }
	uint64_t u64() { uint64_t v = 0; raw(&v, 1); return v; }
	// std::string / std::vector carry a uint64 element count (Interface.h:316-354)
	void str(std::string& s) { const uint64_t n = u64(); if (!ok || n > (1ull<<24)) { ok = false; return; } s.resize((size_t)n); if (n) raw(&s[0], (size_t)n); }
	size_t count(size_t elemBytes) { // a corrupt length must not turn into a huge allocation
		const uint64_t n = u64();
		if (!ok) return 0;
		const long pos = ftell(f); fseek(f, 0, SEEK_END); const long end = ftell(f); fseek(f, pos, SEEK_SET);
		if (end < pos || n > (uint64_t)(end-pos)/elemBytes) { ok = false; return 0; } // division: n*elemBytes may overflow
		return (size_t)n;
	}
};
struct Writer {
	FILE* f;
	template<typename T> void raw(const T* p, size_t n) { if (n) fwrite(p, sizeof(T), n, f); }
	void u32(uint32_t v) { raw(&v, 1); }
	void u64(uint64_t v) { raw(&v, 1); }
	void str(const std::string& s) { u64(s.size()); raw(s.data(), s.size()); }
};

void LoadViews(Reader& r, std::vector<MvsiView>& views) {
	const size_t n = r.count(8); views.resize(n);
	for (MvsiView& v: views) { v.imageID = r.u32(); r.raw(&v.confidence, 1); }
}
void SaveViews(Writer& w, const std::vector<MvsiView>& views) {
	w.u64(views.size());
	for (const MvsiView& v: views) { w.u32(v.imageID); w.raw(&v.confidence, 1); }
}
} // namespace

bool LoadMVSI(const std::string& fileName, MvsiData& obj, uint32_t* pVersion) {
	FILE* f = fopen(fileName.c_str(), "rb");
	if (!f) return false;
	Reader r{f};
	uint32_t version = 0;
	char hdr[4] = {0, 0, 0, 0};
	r.raw(hdr, 4);
	if (!r.ok) { fclose(f); return false; }
	if (strncmp(hdr, "MVSI", 4) != 0) {
		// the first format had no header; only accepted for ".mvs" files (Interface.h:245-254)
		std::string ext = fileName.size() > 4 ? fileName.substr(fileName.size()-4) : std::string();
		std::transform(ext.begin(), ext.end(), ext.begin(), [](char c) { return (char)std::tolower((unsigned char)c); });
		if (ext != ".mvs") { fclose(f); return false; }
		fseek(f, 0, SEEK_SET);
	} else {
		version = r.u32();
		if (!r.ok || version > 5) { fclose(f); return false; }
		r.u32(); // reserved
	}
	obj = MvsiData();
	obj.platforms.resize(r.count(24));
	for (MvsiPlatform& p: obj.platforms) {
		r.str(p.name);
		p.cameras.resize(r.count(8+9*8*2+24));
		for (MvsiCamera& c: p.cameras) {
			r.str(c.name);
			if (version > 3) r.str(c.bandName);
			if (version > 0) { c.width = r.u32(); c.height = r.u32(); }
			r.raw(c.K, 9); r.raw(c.R, 9); r.raw(c.C, 3);
		}
		p.poses.resize(r.count(96));
		for (MvsiPose& q: p.poses) { r.raw(q.R, 9); r.raw(q.C, 3); }
	}
	obj.images.resize(r.count(8+12));
	for (MvsiImage& im: obj.images) {
		r.str(im.name);
		if (version > 4) r.str(im.maskName);
		im.platformID = r.u32(); im.cameraID = r.u32(); im.poseID = r.u32();
		if (version > 2) im.ID = r.u32();
	}
	obj.vertices.resize(r.count(12+8));
	for (MvsiVertex& v: obj.vertices) { r.raw(v.X, 3); LoadViews(r, v.views); }
	obj.verticesNormal.resize(r.count(12)*3); r.raw(obj.verticesNormal.data(), obj.verticesNormal.size());
	obj.verticesColor.resize(r.count(3)*3); r.raw(obj.verticesColor.data(), obj.verticesColor.size());
	if (version > 0) {
		obj.lines.resize(r.count(24+8));
		for (MvsiLine& l: obj.lines) { r.raw(l.pt1, 3); r.raw(l.pt2, 3); LoadViews(r, l.views); }
		obj.linesNormal.resize(r.count(12)*3); r.raw(obj.linesNormal.data(), obj.linesNormal.size());
		obj.linesColor.resize(r.count(3)*3); r.raw(obj.linesColor.data(), obj.linesColor.size());
		if (version > 1) r.raw(obj.transform, 16);
	}
	fclose(f);
	if (pVersion) *pVersion = version;
	return r.ok;
}

bool SaveMVSI(const std::string& fileName, const MvsiData& obj, uint32_t version) {
	if (version > 5) return false;
	FILE* f = fopen(fileName.c_str(), "wb");
	if (!f) return false;
	Writer w{f};
	if (version > 0) { w.raw("MVSI", 4); w.u32(version); w.u32(0); }
	w.u64(obj.platforms.size());
	for (const MvsiPlatform& p: obj.platforms) {
		w.str(p.name);
		w.u64(p.cameras.size());
		for (const MvsiCamera& c: p.cameras) {
			w.str(c.name);
			if (version > 3) w.str(c.bandName);
			if (version > 0) { w.u32(c.width); w.u32(c.height); }
			w.raw(c.K, 9); w.raw(c.R, 9); w.raw(c.C, 3);
		}
		w.u64(p.poses.size());
		for (const MvsiPose& q: p.poses) { w.raw(q.R, 9); w.raw(q.C, 3); }
	}
	w.u64(obj.images.size());
	for (const MvsiImage& im: obj.images) {
		w.str(im.name);
		if (version > 4) w.str(im.maskName);
		w.u32(im.platformID); w.u32(im.cameraID); w.u32(im.poseID);
		if (version > 2) w.u32(im.ID);
	}
	w.u64(obj.vertices.size());
	for (const MvsiVertex& v: obj.vertices) { w.raw(v.X, 3); SaveViews(w, v.views); }
	w.u64(obj.verticesNormal.size()/3); w.raw(obj.verticesNormal.data(), obj.verticesNormal.size());
	w.u64(obj.verticesColor.size()/3); w.raw(obj.verticesColor.data(), obj.verticesColor.size());
	if (version > 0) {
		w.u64(obj.lines.size());
		for (const MvsiLine& l: obj.lines) { w.raw(l.pt1, 3); w.raw(l.pt2, 3); SaveViews(w, l.views); }
		w.u64(obj.linesNormal.size()/3); w.raw(obj.linesNormal.data(), obj.linesNormal.size());
		w.u64(obj.linesColor.size()/3); w.raw(obj.linesColor.data(), obj.linesColor.size());
		if (version > 1) w.raw(obj.transform, 16);
	}
	const bool ok = ferror(f) == 0;
	fclose(f);
	return ok;
}

// ------------------------------------------------------------------------------------------------ cameras
bool ComposeImageCamera(const MvsiData& obj, uint32_t idxImage, uint32_t w, uint32_t h, double K[9], double R[9], double C[3], double Knorm[4]) {
	if (idxImage >= obj.images.size() || !w || !h) return false;
	const MvsiImage& im = obj.images[idxImage];
	if (im.poseID == 0xFFFFFFFFu || im.platformID >= obj.platforms.size()) return false;
	const MvsiPlatform& pl = obj.platforms[im.platformID];
	if (im.cameraID >= pl.cameras.size() || im.poseID >= pl.poses.size()) return false;
	const MvsiCamera& cam = pl.cameras[im.cameraID];
	const MvsiPose& pose = pl.poses[im.poseID];
	// Scene::LoadInterface, Scene.cpp:80-88: K of a camera that carries a resolution is normalised on load
	double Kn[9]; memcpy(Kn, cam.K, sizeof(Kn));
	if (cam.HasResolution()) {
		const double scale = 1.0/(double)(float)std::max(cam.width, cam.height); // REAL(1)/GetNormalizationScale() (a float, Camera.h:105-108)
		Kn[0] *= scale; Kn[4] *= scale; Kn[2] *= scale; Kn[5] *= scale;
	}
	if (Knorm) { Knorm[0] = Kn[0]; Knorm[1] = Kn[4]; Knorm[2] = Kn[2]; Knorm[3] = Kn[5]; }
	// Platform::GetCamera, Platform.cpp:44-54: R = camera.R*pose.R; C = pose.R^T*camera.C + pose.C (cv::Matx products, k ascending)
	for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) {
		double s = 0; for (int k=0; k<3; ++k) s += cam.R[i*3+k]*pose.R[k*3+j];
		R[i*3+j] = s;
	}
	for (int i=0; i<3; ++i) {
		double s = 0; for (int k=0; k<3; ++k) s += pose.R[k*3+i]*cam.C[k];
		C[i] = s+pose.C[i];
	}
	// CameraIntern::GetK<REAL>(width, height), Camera.h:167-180 (identity base: the skew entry is dropped)
	const float fScale = (float)std::max(w, h);
	for (int i=0; i<9; ++i) K[i] = (i%4 == 0) ? 1.0 : 0.0;
	K[0] = Kn[0]*fScale; K[4] = Kn[4]*fScale;
	if (Kn[2] == 0 && Kn[5] == 0) { K[2] = 0.5*(w-1); K[5] = 0.5*(h-1); } // ComposeK, Camera.h:145-153
	else { K[2] = Kn[2]*fScale; K[5] = Kn[5]*fScale; }
	return true;
}

// ------------------------------------------------------------------------------------------------ image files
namespace {
bool ReadAll(const std::string& fileName, std::vector<uint8_t>& buf, size_t limit = 0) {
	FILE* f = fopen(fileName.c_str(), "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
	if (n < 0) { fclose(f); return false; }
	if (limit && (size_t)n > limit) n = (long)limit;
	buf.resize((size_t)n);
	const bool ok = n == 0 || fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
	fclose(f);
	return ok;
}
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0]<<24)|((uint32_t)p[1]<<16)|((uint32_t)p[2]<<8)|p[3]; }
inline uint32_t le32(const uint8_t* p) { return ((uint32_t)p[3]<<24)|((uint32_t)p[2]<<16)|((uint32_t)p[1]<<8)|p[0]; }
inline uint16_t le16(const uint8_t* p) { return (uint16_t)(((uint16_t)p[1]<<8)|p[0]); }
const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

bool DecodePNG(const std::vector<uint8_t>& file, int& w, int& h, std::vector<uint8_t>& bgr, std::string& err) {
	size_t pos = 8;
	int depth = 0, ctype = 0, interlace = 0;
	std::vector<uint8_t> idat, plte;
	bool gotHdr = false, gotEnd = false;
	while (pos+12 <= file.size() && !gotEnd) {
		const uint32_t len = be32(&file[pos]);
		const uint8_t* tag = &file[pos+4];
		if (pos+12+(size_t)len > file.size()) { err = "truncated PNG chunk"; return false; }
		const uint8_t* data = &file[pos+8];
		if (!memcmp(tag, "IHDR", 4) && len >= 13) { w = (int)be32(data); h = (int)be32(data+4); depth = data[8]; ctype = data[9]; interlace = data[12]; gotHdr = true; }
		else if (!memcmp(tag, "PLTE", 4)) plte.assign(data, data+len);
		else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), data, data+len);
		else if (!memcmp(tag, "IEND", 4)) gotEnd = true;
		pos += 12+(size_t)len;
	}
	if (!gotHdr || w <= 0 || h <= 0) { err = "PNG without IHDR"; return false; }
	if (w > 65535 || h > 65535) { err = "PNG larger than 65535 pixels a side"; return false; }
	if (depth != 8 || interlace != 0) { err = "only 8-bit non-interlaced PNG is supported"; return false; }
	int ch;
	switch (ctype) { case 0: ch = 1; break; case 2: ch = 3; break; case 3: ch = 1; break; case 4: ch = 2; break; case 6: ch = 4; break; default: err = "bad PNG colour type"; return false; }
	const size_t stride = (size_t)w*ch;
	std::vector<uint8_t> raw((stride+1)*(size_t)h);
	uLongf outLen = (uLongf)raw.size();
	if (uncompress(raw.data(), &outLen, idat.data(), (uLong)idat.size()) != Z_OK || outLen != raw.size()) { err = "PNG inflate failed"; return false; }
	// un-filter in place (PNG spec 9.2)
	std::vector<uint8_t> img(stride*(size_t)h);
	for (int y=0; y<h; ++y) {
		const uint8_t ft = raw[(stride+1)*y];
		const uint8_t* src = &raw[(stride+1)*y+1];
		uint8_t* dst = &img[stride*y];
		const uint8_t* up = y ? &img[stride*(y-1)] : nullptr;
		for (size_t i=0; i<stride; ++i) {
			const int a = i >= (size_t)ch ? dst[i-ch] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i-ch] : 0;
			int v = src[i];
			switch (ft) {
			case 0: break;
			case 1: v += a; break;
			case 2: v += b; break;
			case 3: v += (a+b)>>1; break;
			case 4: { const int p = a+b-c, pa = std::abs(p-a), pb = std::abs(p-b), pc = std::abs(p-c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
			default: err = "bad PNG filter"; return false;
			}
			dst[i] = (uint8_t)v;
		}
	}
	bgr.resize((size_t)w*h*3);
	for (size_t i=0, n=(size_t)w*h; i<n; ++i) {
		const uint8_t* p = &img[i*ch];
		uint8_t r, g, b;
		if (ctype == 0 || ctype == 4) r = g = b = p[0];
		else if (ctype == 3) { if ((size_t)p[0]*3+2 >= plte.size()) { err = "PNG palette index out of range"; return false; } r = plte[p[0]*3]; g = plte[p[0]*3+1]; b = plte[p[0]*3+2]; }
		else { r = p[0]; g = p[1]; b = p[2]; }
		bgr[i*3] = b; bgr[i*3+1] = g; bgr[i*3+2] = r;
	}
	return true;
}

bool DecodeBMP(const std::vector<uint8_t>& file, int& w, int& h, std::vector<uint8_t>& bgr, std::string& err) {
	if (file.size() < 54) { err = "truncated BMP"; return false; }
	const uint32_t off = le32(&file[10]);
	const int32_t bw = (int32_t)le32(&file[18]), bh = (int32_t)le32(&file[22]);
	const uint16_t bpp = le16(&file[28]);
	const uint32_t comp = le32(&file[30]);
	if ((bpp != 24 && bpp != 32 && bpp != 8) || (comp != 0 && !(comp == 3 && bpp == 32)) || bw <= 0 || bh == 0) { err = "only uncompressed 8/24/32-bit BMP is supported"; return false; }
	w = bw; h = bh < 0 ? -bh : bh;
	if (w > 65535 || h > 65535) { err = "BMP larger than 65535 pixels a side"; return false; }
	if (bpp == 8 && (size_t)14+le32(&file[14])+1024 > file.size()) { err = "truncated BMP palette"; return false; }
	const size_t rowBytes = (((size_t)w*bpp+31)/32)*4;
	if ((size_t)off+rowBytes*(size_t)h > file.size()) { err = "truncated BMP pixels"; return false; }
	const uint8_t* pal = bpp == 8 ? &file[14+le32(&file[14])] : nullptr;
	bgr.resize((size_t)w*h*3);
	for (int y=0; y<h; ++y) {
		const uint8_t* src = &file[off+rowBytes*(size_t)(bh < 0 ? y : h-1-y)];
		uint8_t* dst = &bgr[(size_t)y*w*3];
		for (int x=0; x<w; ++x) {
			if (bpp == 8) { const uint8_t* e = pal+src[x]*4; dst[x*3] = e[0]; dst[x*3+1] = e[1]; dst[x*3+2] = e[2]; }
			else { const uint8_t* p = src+(size_t)x*(bpp/8); dst[x*3] = p[0]; dst[x*3+1] = p[1]; dst[x*3+2] = p[2]; }
		}
	}
	return true;
}

bool ParsePNMHeader(const std::vector<uint8_t>& file, int& kind, int& w, int& h, int& maxv, size_t& dataPos) {
	if (file.size() < 2 || file[0] != 'P' || (file[1] != '5' && file[1] != '6')) return false;
	kind = file[1]-'0';
	size_t p = 2; int vals[3], nv = 0;
	while (nv < 3 && p < file.size()) {
		while (p < file.size() && std::isspace(file[p])) ++p;
		if (p < file.size() && file[p] == '#') { while (p < file.size() && file[p] != '\n') ++p; continue; }
		int v = 0; bool any = false;
		while (p < file.size() && std::isdigit(file[p])) { v = v*10+(file[p]-'0'); ++p; any = true; }
		if (!any) return false;
		vals[nv++] = v;
	}
	if (nv != 3) return false;
	w = vals[0]; h = vals[1]; maxv = vals[2]; dataPos = p+1; // one whitespace byte after maxval
	return w > 0 && h > 0 && w <= 65535 && h <= 65535;
}
} // namespace

bool LoadImageBGR(const std::string& fileName, int& w, int& h, std::vector<uint8_t>& bgr, std::string* perr) {
	std::string err;
	std::vector<uint8_t> file;
	bool ok = false;
	if (!ReadAll(fileName, file)) err = "cannot read '"+fileName+"'";
	else if (file.size() >= 8 && !memcmp(file.data(), kPngSig, 8)) ok = DecodePNG(file, w, h, bgr, err);
	else if (file.size() >= 2 && file[0] == 'B' && file[1] == 'M') ok = DecodeBMP(file, w, h, bgr, err);
	else {
		int kind, maxv; size_t pos;
		if (ParsePNMHeader(file, kind, w, h, maxv, pos)) {
			const size_t n = (size_t)w*h, need = n*(kind == 6 ? 3 : 1);
			if (maxv != 255 || pos+need > file.size()) err = "only 8-bit binary PGM/PPM is supported";
			else {
				bgr.resize(n*3);
				for (size_t i=0; i<n; ++i) {
					if (kind == 6) { bgr[i*3] = file[pos+i*3+2]; bgr[i*3+1] = file[pos+i*3+1]; bgr[i*3+2] = file[pos+i*3]; }
					else bgr[i*3] = bgr[i*3+1] = bgr[i*3+2] = file[pos+i];
				}
				ok = true;
			}
		} else err = "unsupported image format '"+fileName+"' (BMP, PNG and binary PGM/PPM are read; the other libs/IO codecs are out of scope)";
	}
	if (!ok && perr) *perr = err;
	return ok;
}

bool ReadImageSize(const std::string& fileName, int& w, int& h) {
	std::vector<uint8_t> file;
	if (!ReadAll(fileName, file, 4096) || file.size() < 26) return false;
	if (!memcmp(file.data(), kPngSig, 8) && !memcmp(&file[12], "IHDR", 4)) { w = (int)be32(&file[16]); h = (int)be32(&file[20]); return w > 0 && h > 0; }
	if (file[0] == 'B' && file[1] == 'M') { w = (int32_t)le32(&file[18]); const int32_t bh = (int32_t)le32(&file[22]); h = bh < 0 ? -bh : bh; return w > 0 && h > 0; }
	int kind, maxv; size_t pos;
	return ParsePNMHeader(file, kind, w, h, maxv, pos);
}

// ------------------------------------------------------------------------------------------------ Scene <-> MVSI
namespace {
std::string DirOf(const std::string& path) { const size_t p = path.find_last_of("/\\"); return p == std::string::npos ? std::string() : path.substr(0, p+1); }
bool IsAbs(const std::string& p) { return !p.empty() && (p[0] == '/' || (p.size() > 1 && p[1] == ':')); }
} // namespace

bool Scene::LoadInterface(const std::string& fileName, bool bLoadImages, std::string* perr) {
	// Scene::LoadInterface, libs/MVS/Scene.cpp:62-216
	auto fail = [&](const std::string& m) { if (perr) *perr = m; return false; };
	MvsiData obj;
	if (!LoadMVSI(fileName, obj)) return fail("cannot load MVSI project '"+fileName+"'");
	if (obj.platforms.empty()) return fail("project without platforms");
	const std::string dir = DirOf(fileName); // WORKING_FOLDER: the project's own folder unless the caller changed directory
	images.clear(); pointcloud = SparsePoints();
	images.resize(obj.images.size());
	unsigned nCalibrated = 0;
	for (size_t i=0; i<obj.images.size(); ++i) {
		const MvsiImage& src = obj.images[i];
		Image& im = images[i];
		im.name = src.name;
		std::replace(im.name.begin(), im.name.end(), '\\', '/'); // Util::ensureUnifySlash
		if (!IsAbs(im.name)) im.name = dir+im.name;               // MAKE_PATH_FULL(WORKING_FOLDER_FULL, name)
		im.ID = src.ID == 0xFFFFFFFFu ? (uint32_t)i : src.ID;
		im.calibrated = false;
		if (src.poseID == 0xFFFFFFFFu) continue; // uncalibrated image: stays in the list, takes no part (Scene.cpp:138-141)
		if (src.platformID >= obj.platforms.size() || src.cameraID >= obj.platforms[src.platformID].cameras.size() ||
		    src.poseID >= obj.platforms[src.platformID].poses.size()) return fail("image '"+src.name+"' refers to a missing platform / camera / pose");
		const MvsiCamera& cam = obj.platforms[src.platformID].cameras[src.cameraID];
		int w = (int)cam.width, h = (int)cam.height;
		if (!cam.HasResolution() && !ReadImageSize(im.name, w, h)) return fail("cannot read the header of '"+im.name+"'"); // Image::ReloadImage(0, false)
		if (bLoadImages) {
			int iw, ih; std::string e;
			if (!LoadImageBGR(im.name, iw, ih, im.bgr, &e)) return fail(e);
			if (iw != w || ih != h) return fail("image '"+im.name+"' does not have the resolution its camera states (rescaling on load, Image.cpp:139-160, is not built)");
			im.gray.resize((size_t)w*h);
			ToGray(im.bgr.data(), w, h, im.gray.data());
		}
		im.width = w; im.height = h;
		if (!ComposeImageCamera(obj, (uint32_t)i, (uint32_t)w, (uint32_t)h, im.camera.K, im.camera.R, im.camera.C, im.Knorm)) return fail("bad camera of image '"+src.name+"'");
		im.camera.ComposeP();
		im.hasKnorm = true;
		im.calibrated = true;
		++nCalibrated;
	}
	if (images.size() < 2) return fail("a project needs at least 2 images");
	// 3-D points: views sorted by image id (stable on the original order through the index sort, Scene.cpp:166-178)
	pointcloud.xyz.resize(obj.vertices.size()*3);
	pointcloud.views.resize(obj.vertices.size());
	pointcloud.weights.clear();
	bool validWeights = false;
	std::vector<std::vector<float>> wts(obj.vertices.size());
	for (size_t i=0; i<obj.vertices.size(); ++i) {
		const MvsiVertex& v = obj.vertices[i];
		memcpy(&pointcloud.xyz[i*3], v.X, 12);
		std::vector<uint32_t> idx(v.views.size());
		std::iota(idx.begin(), idx.end(), 0u);
		std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return v.views[a].imageID < v.views[b].imageID; });
		pointcloud.views[i].resize(idx.size()); wts[i].resize(idx.size());
		for (size_t k=0; k<idx.size(); ++k) {
			pointcloud.views[i][k] = v.views[idx[k]].imageID; wts[i][k] = v.views[idx[k]].confidence;
			if (wts[i][k] != 0) validWeights = true;
		}
	}
	if (validWeights) pointcloud.weights.swap(wts);
	pointcloud.normals = obj.verticesNormal;
	pointcloud.colors = obj.verticesColor;
	if (perr) perr->clear();
	return nCalibrated >= 2 || fail("fewer than 2 calibrated images");
}

bool Scene::SaveInterface(const std::string& fileName, int version, bool bDense) const {
	// Scene::SaveInterface, libs/MVS/Scene.cpp:218-286. This mirror keeps absolute cameras per image, so every image gets its own
	// camera + pose on one platform: full-resolution K with the image size (Interface.h:381-383, the reference loads both this and the
	// normalised form) and identity relative pose, pose = (R, C) — what LoadInterface composes back to the same absolute camera.
	MvsiData obj;
	obj.platforms.resize(1);
	MvsiPlatform& pl = obj.platforms[0];
	const std::string dir = DirOf(fileName);
	obj.images.resize(images.size());
	for (size_t i=0; i<images.size(); ++i) {
		const Image& im = images[i];
		MvsiImage& dst = obj.images[i];
		dst.name = (!dir.empty() && im.name.compare(0, dir.size(), dir) == 0) ? im.name.substr(dir.size()) : im.name; // MAKE_PATH_REL
		dst.ID = im.ID == 0xFFFFFFFFu ? (uint32_t)i : im.ID;
		if (!im.calibrated && !(im.width > 0 && im.height > 0)) continue;
		MvsiCamera cam;
		cam.width = (uint32_t)im.width; cam.height = (uint32_t)im.height; // the camera carries its resolution: no image header is needed to load it back
		memcpy(cam.K, im.camera.K, sizeof(cam.K));
		MvsiPose pose; memcpy(pose.R, im.camera.R, sizeof(pose.R)); memcpy(pose.C, im.camera.C, sizeof(pose.C));
		dst.platformID = 0; dst.cameraID = (uint32_t)pl.cameras.size(); dst.poseID = (uint32_t)pl.poses.size();
		pl.cameras.push_back(cam); pl.poses.push_back(pose);
	}
	if (bDense && densecloud.size()) {
		const size_t n = densecloud.size();
		obj.vertices.resize(n);
		for (size_t i=0; i<n; ++i) {
			MvsiVertex& v = obj.vertices[i];
			memcpy(v.X, &densecloud.points[i*3], 12);
			const uint32_t a = densecloud.viewOffsets[i], b = densecloud.viewOffsets[i+1];
			v.views.resize(b-a);
			for (uint32_t k=a; k<b; ++k) { v.views[k-a].imageID = densecloud.views[k]; v.views[k-a].confidence = densecloud.weights.empty() ? 0.f : densecloud.weights[k]; }
		}
		if (!densecloud.normals.empty()) obj.verticesNormal.assign(densecloud.normals.data(), densecloud.normals.data()+n*3);
		if (!densecloud.colors.empty()) obj.verticesColor.assign(densecloud.colors.data(), densecloud.colors.data()+n*3);
	} else {
		const size_t n = pointcloud.size();
		obj.vertices.resize(n);
		for (size_t i=0; i<n; ++i) {
			MvsiVertex& v = obj.vertices[i];
			memcpy(v.X, &pointcloud.xyz[i*3], 12);
			v.views.resize(pointcloud.views[i].size());
			for (size_t k=0; k<v.views.size(); ++k) { v.views[k].imageID = pointcloud.views[i][k]; v.views[k].confidence = pointcloud.weights.empty() ? 0.f : pointcloud.weights[i][k]; }
		}
		obj.verticesNormal = pointcloud.normals;
		obj.verticesColor = pointcloud.colors;
	}
	return SaveMVSI(fileName, obj, version >= 0 ? (uint32_t)version : 5u);
}

} // namespace hcmvs_host
