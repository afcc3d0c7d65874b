// The fork's per-map hand-off files between the pyramid levels of run.sh: depthmap/depthNNNN.dmap and normalmap/normalNNNN.dmap
// (MVS::SaveDepthMap / LoadDepthMap / SaveNormalMap / LoadNormalMap, libs/MVS/DepthMap.cpp:2368-2393; written at
// SceneDensify.cpp:3984-3989): SerializeSave(map, file, ARCHIVE_BINARY_ZIP) = a Boost.Serialization binary_oarchive of the map pushed
// through boost::iostreams::zlib_compressor(best_speed) (libs/Common/Types.inl:3753-3760).
//
// Boost is not available here, so the archive is written / parsed from its documented binary layout (boost/archive/basic_binary_oarchive,
// basic_binary_oprimitive, detail/oserializer: no class ids in binary archives; the first object of every class carries tracking_type
// (1 byte) + version_type (4 bytes, library version > 7)). BYTE PARITY WITH A REAL BOOST BUILD IS NOT VERIFIED (no Boost to check
// against); the tests round-trip through this reader and re-derive the layout independently in Python (zlib + struct):
//
//   u64 22, "serialization::archive", u16 library version (17 = Boost 1.71, the Dockerfile's Ubuntu 20.04)   basic_binary_oarchive::init
//   u8 sizeof(int) 4, u8 sizeof(long) 8, u8 sizeof(float) 4, u8 sizeof(double) 8, i32 1 (endianness probe)      basic_binary_oprimitive::init
//   DepthMap = TImage<float> : TDMatrix<float> : cv::Mat_<float>, each level `ar & base_object<Base>(*this)` (Common/Types.h:2094-2180):
//     3 x (u8 tracking 0, u32 version 0), i32 cols, i32 rows, cols*rows floats (make_array of an arithmetic type: one raw block)
//   NormalMap = TImage<Point3f> ...: the same 3 levels, i32 cols, i32 rows, then the elements ONE BY ONE (Point3f is a class, so
//     make_array is not "bitwise": array_wrapper falls back to `ar & item`): the first element carries the class info of
//     TPoint3<float> and of its base cv::Point3_<float> (2 x 5 bytes), every element is x, y, z (Types.inl:3671-3675).
#include "densify.h"
#include <zlib.h>
#include <cstdio>
#include <cstring>

namespace hcmvs_host {

namespace {
const char kSignature[] = "serialization::archive";
const uint16_t kLibraryVersion = 17; // Boost 1.71

struct Buf {
	std::vector<unsigned char> d; size_t pos = 0; bool bad = false;
	template<typename T> void put(T v) { const unsigned char* p = (const unsigned char*)&v; d.insert(d.end(), p, p+sizeof(T)); }
	void put(const void* p, size_t n) { const unsigned char* q = (const unsigned char*)p; d.insert(d.end(), q, q+n); }
	template<typename T> T get() { T v = T(); if (pos+sizeof(T) > d.size()) { bad = true; return v; } std::memcpy(&v, d.data()+pos, sizeof(T)); pos += sizeof(T); return v; }
	bool get(void* p, size_t n) { if (pos+n > d.size()) { bad = true; return false; } std::memcpy(p, d.data()+pos, n); pos += n; return true; }
};

void WriteHeader(Buf& b) {
	b.put<uint64_t>(sizeof(kSignature)-1); b.put(kSignature, sizeof(kSignature)-1); b.put<uint16_t>(kLibraryVersion);
	b.put<uint8_t>(4); b.put<uint8_t>(8); b.put<uint8_t>(4); b.put<uint8_t>(8); b.put<int32_t>(1);
}
bool ReadHeader(Buf& b, uint16_t& lib) {
	const uint64_t n = b.get<uint64_t>();
	char sig[32] = {0};
	if (b.bad || n != sizeof(kSignature)-1 || !b.get(sig, (size_t)n) || std::memcmp(sig, kSignature, (size_t)n) != 0) return false;
	lib = b.get<uint16_t>();
	const uint8_t si = b.get<uint8_t>(), sl = b.get<uint8_t>(), sf = b.get<uint8_t>(), sd = b.get<uint8_t>();
	const int32_t endian = b.get<int32_t>();
	return !b.bad && lib > 7 && si == 4 && sl == 8 && sf == 4 && sd == 8 && endian == 1; // what a 64-bit little-endian Linux build writes
}
void WriteClassInfo(Buf& b, int levels) { for (int i=0; i<levels; ++i) { b.put<uint8_t>(0); b.put<uint32_t>(0); } } // tracking_type, version_type
bool ReadClassInfo(Buf& b, int levels) { for (int i=0; i<levels; ++i) { b.get<uint8_t>(); b.get<uint32_t>(); } return !b.bad; }

bool Deflate(const Buf& b, const std::string& fileName) { // io::zlib_compressor(io::zlib::best_speed): a zlib (RFC 1950) stream, level 1
	uLongf cap = compressBound((uLong)b.d.size());
	std::vector<unsigned char> z(cap);
	if (compress2(z.data(), &cap, b.d.data(), (uLong)b.d.size(), Z_BEST_SPEED) != Z_OK) return false;
	FILE* f = fopen(fileName.c_str(), "wb");
	if (!f) return false;
	const bool ok = fwrite(z.data(), 1, cap, f) == cap;
	return fclose(f) == 0 && ok;
}
bool Inflate(const std::string& fileName, Buf& b) {
	FILE* f = fopen(fileName.c_str(), "rb");
	if (!f) return false;
	std::vector<unsigned char> z; unsigned char tmp[1<<16]; size_t n;
	while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) z.insert(z.end(), tmp, tmp+n);
	fclose(f);
	z_stream s; std::memset(&s, 0, sizeof(s));
	if (inflateInit(&s) != Z_OK) return false;
	s.next_in = z.data(); s.avail_in = (uInt)z.size();
	int r = Z_OK;
	while (r == Z_OK) {
		const size_t at = b.d.size();
		b.d.resize(at+(1<<20));
		s.next_out = b.d.data()+at; s.avail_out = 1<<20;
		r = inflate(&s, Z_NO_FLUSH);
		b.d.resize(at+((1<<20)-s.avail_out));
		if (r == Z_OK && s.avail_in == 0 && s.avail_out != 0) break; // truncated stream
	}
	inflateEnd(&s);
	return r == Z_STREAM_END;
}
} // namespace

bool SaveDepthMap(const std::string& fileName, const float* depth, int w, int h) {
	if (!depth || w <= 0 || h <= 0) return false;
	Buf b; WriteHeader(b); WriteClassInfo(b, 3);
	b.put<int32_t>(w); b.put<int32_t>(h);
	b.put(depth, (size_t)w*h*4);
	return Deflate(b, fileName);
}
bool LoadDepthMap(const std::string& fileName, std::vector<float>& depth, int& w, int& h) {
	Buf b; uint16_t lib;
	if (!Inflate(fileName, b) || !ReadHeader(b, lib) || !ReadClassInfo(b, 3)) return false;
	w = b.get<int32_t>(); h = b.get<int32_t>();
	if (b.bad || w <= 0 || h <= 0 || (uint64_t)w*(uint64_t)h > ((uint64_t)1<<32)) return false;
	depth.resize((size_t)w*h);
	return b.get(depth.data(), depth.size()*4) && b.pos == b.d.size();
}
bool SaveNormalMap(const std::string& fileName, const float* normal, int w, int h) {
	if (!normal || w <= 0 || h <= 0) return false;
	Buf b; WriteHeader(b); WriteClassInfo(b, 3);
	b.put<int32_t>(w); b.put<int32_t>(h);
	WriteClassInfo(b, 2); // the first element's TPoint3<float> and cv::Point3_<float>
	b.put(normal, (size_t)w*h*12);
	return Deflate(b, fileName);
}
bool LoadNormalMap(const std::string& fileName, std::vector<float>& normal, int& w, int& h) {
	Buf b; uint16_t lib;
	if (!Inflate(fileName, b) || !ReadHeader(b, lib) || !ReadClassInfo(b, 3)) return false;
	w = b.get<int32_t>(); h = b.get<int32_t>();
	if (b.bad || w <= 0 || h <= 0 || (uint64_t)w*(uint64_t)h > ((uint64_t)1<<32)) return false;
	if (!ReadClassInfo(b, 2)) return false;
	normal.resize((size_t)w*h*3);
	return b.get(normal.data(), normal.size()*4) && b.pos == b.d.size();
}

} // namespace hcmvs_host
