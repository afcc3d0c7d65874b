// MVSI project files (".mvs") and the image files they name — the INPUT side of the drop-in boundary (SURVEY §8b "Inputs").
//
//   reference                                                   here
//   MVS::Interface + ARCHIVE (libs/MVS/Interface.h:165-619)     hcmvs_host::MvsiData, LoadMVSI / SaveMVSI
//   Scene::LoadInterface / SaveInterface (Scene.cpp:62-286)     Scene::LoadInterface / Scene::SaveInterface
//   Platform::GetCamera (Platform.cpp:44-54), Image::GetCamera
//   (Image.cpp:194-209), CameraIntern::GetK (Camera.h:167-180)  ComposeImageCamera
//
// The byte layout is pinned against the reference's own header: oracle/_ref/mvsi_ref_tool is compiled from
// /root/reference/frame_main/libs/MVS/Interface.h where it lies and tests/test_mvsi_ref.py cross-reads / cross-writes files
// with it (byte-identical output).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace hcmvs_host {

struct MvsiCamera { // Interface::Platform::Camera, Interface.h:375-404
	std::string name, bandName;
	uint32_t width = 0, height = 0; // 0 = K is normalised by max(width, height) of the image
	double K[9] = {1,0,0, 0,1,0, 0,0,1}, R[9] = {1,0,0, 0,1,0, 0,0,1}, C[3] = {0,0,0};
	bool HasResolution() const { return width > 0 && height > 0; }
};
struct MvsiPose { double R[9] = {1,0,0, 0,1,0, 0,0,1}, C[3] = {0,0,0}; }; // Interface::Platform::Pose, :407-425
struct MvsiPlatform { std::string name; std::vector<MvsiCamera> cameras; std::vector<MvsiPose> poses; }; // :373-468
struct MvsiImage { // Interface::Image, :472-499
	std::string name, maskName;
	uint32_t platformID = 0xFFFFFFFFu, cameraID = 0xFFFFFFFFu, poseID = 0xFFFFFFFFu, ID = 0xFFFFFFFFu;
};
struct MvsiView { uint32_t imageID; float confidence; }; // Interface::Vertex::View, :504-515
struct MvsiVertex { float X[3]; std::vector<MvsiView> views; }; // :502-526
struct MvsiLine { float pt1[3], pt2[3]; std::vector<MvsiView> views; }; // :529-555

struct MvsiData { // Interface, :363-619
	std::vector<MvsiPlatform> platforms;
	std::vector<MvsiImage> images;
	std::vector<MvsiVertex> vertices;
	std::vector<float> verticesNormal;   // 3 per entry
	std::vector<uint8_t> verticesColor;  // 3 per entry, stored B,G,R (Col3: x=B y=G z=R)
	std::vector<MvsiLine> lines;
	std::vector<float> linesNormal;
	std::vector<uint8_t> linesColor;
	double transform[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1};
};

// ARCHIVE::SerializeLoad / SerializeSave, Interface.h:211-270. version: 0 = header-less first format (".mvs" only) .. 5.
bool LoadMVSI(const std::string& fileName, MvsiData& obj, uint32_t* pVersion = nullptr);
bool SaveMVSI(const std::string& fileName, const MvsiData& obj, uint32_t version = 5);

// absolute, un-normalised camera of an image at resolution w x h, exactly as Scene::LoadInterface + Image::UpdateCamera build it
// (K normalisation Scene.cpp:80-88, pose composition Platform.cpp:44-54, CameraIntern::GetK Camera.h:167-180)
bool ComposeImageCamera(const MvsiData& obj, uint32_t idxImage, uint32_t w, uint32_t h, double K[9], double R[9], double C[3], double Knorm[4] = nullptr);

// 8-bit image files -> BGR. BMP (24/32-bit BI_RGB), PNG (8-bit gray / RGB / RGBA / palette, non-interlaced; zlib) and binary
// PPM/PGM. JPG/TIFF/DDS/TGA/SCI (libs/IO) are out of scope: a clear error, never a silent fallback.
bool LoadImageBGR(const std::string& fileName, int& w, int& h, std::vector<uint8_t>& bgr, std::string* err = nullptr);
bool ReadImageSize(const std::string& fileName, int& w, int& h);

} // namespace hcmvs_host
