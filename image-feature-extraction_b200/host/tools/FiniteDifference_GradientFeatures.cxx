// FiniteDifference_GradientFeatures -i/--image -m/--mask -o/--outdir [-p/--prefix gradient_]
// Flags, semantics and output name of the reference tool
// (tools/FiniteDifference_GradientFeatures.cxx:31-68,105-124): gradient magnitude of the raw
// image masked by a mask that is read as float; output
//   <outdir>/<prefix>GradientMagnitude.nii.gz
#include <cstdlib>
#include <iostream>
#include <string>

#include "ife/Context.h"
#include "ife/IO/NiftiIO.h"
#include "ife/Util/CmdLine.h"
#include "ife/Util/Path.h"

const std::string VERSION("0.1");
const std::string OUT_FILE_TYPE(".nii.gz");

int main(int argc, char* argv[]) {
  ife::CmdLine cmd("Calculate gradient based features.", VERSION);
  cmd.add("i", "image", "Path to image.", true, "", "path");
  cmd.add("m", "mask", "Path to mask. Must match image dimensions.", true, "", "path");
  cmd.add("o", "outdir", "Path to output directory", true, "", "path");
  cmd.add("p", "prefix", "Prefix to use for output filenames", false, "gradient_", "string");
  int rc;
  if (!cmd.parse(argc, argv, &rc)) return rc;
  const std::string imagePath(cmd.value("image")), maskPath(cmd.value("mask"));
  const std::string baseFileName = ife::Path::join(cmd.value("outdir"), cmd.value("prefix"));
  try {
    auto both = ife::nifti::ReadPair<float, float>(imagePath, maskPath);   // the two files are inflated concurrently
    ife::Image<float>::Pointer image = both.first;
    ife::Image<float>::Pointer mask = both.second;
    if (mask->GetSize() != image->GetSize()) throw std::runtime_error("mask and image dimensions differ");
    auto out = ife::Image<float>::New();
    out->SetGeometry(image->GetGeometry());
    out->Allocate();
    const ife::Geometry& g = image->GetGeometry();
    ife::CudaContext& c = ife::CudaContext::Instance();
    c.Check(ife_cuda_gradient_magnitude(c.Handle(), image->GetBufferPointer(), mask->GetBufferPointer(), nullptr,
                                        out->GetBufferPointer(), g.size.data(), g.spacing.data(), IFE_MEM_HOST));
    ife::nifti::Write(baseFileName + "GradientMagnitude" + OUT_FILE_TYPE, *out);
  } catch (std::exception& e) {
    std::cerr << "Failed to process." << std::endl
              << "Image: " << imagePath << std::endl
              << "Mask: " << maskPath << std::endl
              << "Base file name: " << baseFileName << std::endl
              << "ExceptionObject: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
