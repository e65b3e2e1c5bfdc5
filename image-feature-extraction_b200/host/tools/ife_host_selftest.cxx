// Host-side self test that needs no GPU: command-line parsing, Path::join, NIfTI round trip.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include <vector>

#include "ife/IO/NiftiIO.h"
#include "ife/IO/ROIReader.h"
#include "ife/Statistics/DetermineEdgesForEqualizedHistogram.h"
#include "ife/Util/CmdLine.h"
#include "ife/Util/Path.h"

#define CHECK(c) do { if (!(c)) { std::cerr << "FAILED: " #c << " (line " << __LINE__ << ")" << std::endl; return 1; } } while (0)

int main(int argc, char* argv[]) {
  CHECK(ife::Path::join("a/b//", "/c.nii") == "a/b/c.nii");
  CHECK(ife::Path::join("", "x") == "/x");
  CHECK(ife::Path::join("out", "") == "out/");
  {
    ife::CmdLine cmd("t", "0.1");
    cmd.add("i", "image", "", true, "", "path");
    cmd.add("s", "scale", "", true, "", "double", true);
    cmd.add("p", "prefix", "", false, "def_", "string");
    const char* av[] = {"tool", "-i", "a.nii", "--scale", "0.6", "-s", "1.2", "--scale=2.4"};
    int rc = -1;
    CHECK(cmd.parse(8, const_cast<char**>(av), &rc));
    CHECK(cmd.value("image") == "a.nii" && cmd.values("scale").size() == 3 && cmd.value("prefix") == "def_");
    float f;
    CHECK(ife::CmdLine::convert(cmd.values("scale")[2], &f) && f == 2.4f);
    CHECK(std::to_string(0.6f) == "0.600000");
  }
  {  // test/DetermineEdgesForEqualizedHistogramTest.cxx:30-70
    std::vector<double> v{1, 2, 3, 4, 5, 6, 7, 8, 9}, e(2);
    ife::determineEdgesForEqualizedHistogram(v.begin(), v.end(), e.begin(), 3);
    CHECK(e[0] == 4 && e[1] == 7);
    std::vector<double> ones(8, 1), e1{0, 123};
    ife::determineEdgesForEqualizedHistogram(ones.begin(), ones.end(), e1.begin(), 2);
    CHECK(e1[0] == 1);
    std::vector<double> u{1, 1, 1, 1, 1, 2, 2, 3, 3, 3};
    ife::determineEdgesForEqualizedHistogram(u.begin(), u.end(), e.begin(), 3);
    CHECK(e[0] == 2 && e[1] == 3);
    bool threw = false;
    std::vector<double> e9(9);
    try { ife::determineEdgesForEqualizedHistogram(v.begin(), v.end(), e9.begin(), 10); } catch (const std::out_of_range&) { threw = true; }
    CHECK(threw);
  }
  {  // ROI text format round trip
    std::stringstream ss;
    ife::ROIReader::write(ss, {ife::Region{{1, 2, 3, 41, 41, 41}}, ife::Region{{0, 0, 7, 5, 6, 7}}});
    CHECK(ss.str() == "[1, 2, 3][41, 41, 41]\n[0, 0, 7][5, 6, 7]\n");
    auto back = ife::ROIReader::read(ss, false);
    CHECK(back.size() == 2 && back[1][2] == 7 && back[0][5] == 41);
  }
  const std::string dir = argc > 1 ? argv[1] : "/tmp";
  auto img = ife::Image<float>::New();
  img->SetRegions(5, 4, 3);
  img->SetSpacing(0.7, 0.8, 2.5);
  img->Allocate();
  for (size_t i = 0; i < img->GetNumberOfPixels(); ++i) img->GetBufferPointer()[i] = (float)i * 0.5f - 7;
  for (const char* ext : {".nii", ".nii.gz"}) {
    const std::string p = ife::Path::join(dir, std::string("ife_selftest") + ext);
    ife::nifti::Write(p, *img);
    auto back = ife::nifti::Read<float>(p);
    CHECK(back->GetSize() == img->GetSize());
    CHECK(std::fabs(back->GetSpacing()[2] - 2.5) < 1e-6 && std::fabs(back->GetSpacing()[0] - 0.7) < 1e-6);
    CHECK(back->SameBufferContent(*img));
    auto as_u8 = ife::nifti::Read<unsigned char>(p);   // cast on read, like itk::ImageFileReader
    CHECK(as_u8->GetPixel(4, 3, 2) == (unsigned char)(59 * 0.5f - 7));
    std::remove(p.c_str());
  }
  {
    // a .nii.gz of several 8 MB pieces (written by several threads as ONE gzip member), with a
    // ragged last piece; read back through zlib's ordinary single-stream reader
    auto big = ife::Image<float>::New();
    big->SetRegions(256, 160, 131);
    big->Allocate();
    uint32_t lcg = 12345u;
    for (size_t i = 0; i < big->GetNumberOfPixels(); ++i) {
      lcg = lcg * 1664525u + 1013904223u;
      big->GetBufferPointer()[i] = (i % 4096 < 1024) ? 0.0f : (float)(lcg >> 8) * (1.0f / 65536.0f) - 100.0f;
    }
    for (const char* threads : {"5", "1"}) {
      setenv("IFE_IO_THREADS", threads, 1);
      const std::string p = ife::Path::join(dir, "ife_selftest_big.nii.gz");
      ife::nifti::Write(p, *big);
      auto back = ife::nifti::Read<float>(p);
      CHECK(back->GetSize() == big->GetSize());
      CHECK(back->SameBufferContent(*big));
      if (argc > 2) break;   // keep the file for the caller (tests read it with Python's gzip)
      std::remove(p.c_str());
    }
    unsetenv("IFE_IO_THREADS");
  }
  std::cout << "host selftest ok" << std::endl;
  return 0;
}
