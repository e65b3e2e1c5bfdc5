// MakeBagDense -i image -m mask -H histogram-spec -o outdir -s scale ...
//              [-M roi-mask] [-v roi-mask-value] [-x/-y/-z roi-size] [-p prefix]
// Flags, file formats and semantics of the reference tool (tools/MakeBagDense.cxx): ONE ROI per
// non-zero voxel of the ROI mask (default: the clamped image mask), start = index - size/2,
// kept when it lies inside the image (include/ife/ROI/DenseROIGenerator.hxx:22-45), written to
// <outdir>/<prefix>.ROIInfo (:265-273); per ROI every in-mask voxel's 8 features at every scale
// are inserted into DenseHistograms and the frequencies become one CSV row of
// <outdir>/<prefix>.bag (:381-432).  On the GPU: the fused kernel leaves the eight bin indices
// of every voxel in one word and one thread block per ROI counts its box
// (ife_cuda_emphysema_histograms with a long ROI list); ROIs are processed in slices of 32 Ki
// so that the counts of a slice (5 KB per ROI at 4 scales x 41 bins) stay small.
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "ife/Context.h"
#include "ife/IO/NiftiIO.h"
#include "ife/IO/ROIReader.h"
#include "ife/Util/CmdLine.h"
#include "ife/Util/Path.h"

const std::string VERSION("0.1");
const size_t numFeatures = 8;

int main(int argc, char* argv[]) {
  ife::CmdLine cmd("Create a bag of instances samples from an image.", VERSION);
  cmd.add("i", "image", "Path to image.", true, "", "path");
  cmd.add("m", "mask", "Path to mask.", true, "", "path");
  cmd.add("H", "histogram-spec", "Path to histogram specification.", true, "", "path");
  cmd.add("o", "outdir", "Path to output directory", true, "", "path");
  cmd.add("s", "scale", "Scales for the Gauss applicability function", true, "", "double", true);
  cmd.add("M", "roi-mask", "Path to ROI mask file. If ROIs are generated an optional mask controlling the ROI generation can be used. If not given then the image mask will be used.", false, "", "path");
  cmd.add("v", "roi-mask-value", "Value in the ROI mask that should be used for inclusion.", false, "1", "MaskPixelType");
  cmd.add("x", "roi-size-x", "Size of ROI in x dimension", false, "41", "N>=1");
  cmd.add("y", "roi-size-y", "Size of ROI in y dimension", false, "41", "N>=1");
  cmd.add("z", "roi-size-z", "Size of ROI in z dimension", false, "41", "N>=1");
  cmd.add("p", "prefix", "Prefix to use for output filenames", false, "", "string");
  int rc;
  if (!cmd.parse(argc, argv, &rc)) return rc;
  const std::string imagePath(cmd.value("image")), maskPath(cmd.value("mask")), histPath(cmd.value("histogram-spec"));
  const std::string outDirPath(cmd.value("outdir")), roiMaskPath(cmd.value("roi-mask"));
  const std::string prefix(cmd.value("prefix"));
  std::vector<double> scales;
  for (const std::string& s : cmd.values("scale")) {
    float v;
    if (!ife::CmdLine::convert(s, &v)) { cmd.error("Couldn't read argument value from string '" + s + "'", "-s", &rc); return rc; }
    scales.push_back(v);
  }
  size_t roiSize[3] = {41, 41, 41};
  unsigned roiMaskValue = 1;
  if (!ife::CmdLine::convert(cmd.value("roi-size-x"), &roiSize[0]) ||
      !ife::CmdLine::convert(cmd.value("roi-size-y"), &roiSize[1]) || !ife::CmdLine::convert(cmd.value("roi-size-z"), &roiSize[2]) ||
      !ife::CmdLine::convert(cmd.value("roi-mask-value"), &roiMaskValue)) {
    cmd.error("Couldn't read a numeric argument value", "", &rc);
    return rc;
  }

  try {
    auto both = ife::nifti::ReadPair<float, unsigned short>(imagePath, maskPath);   // the two files are inflated concurrently
    ife::Image<float>::Pointer image = both.first;
    ife::Image<unsigned short>::Pointer mask16 = both.second;
    const ife::Geometry& g = image->GetGeometry();
    if (mask16->GetSize() != image->GetSize()) throw std::runtime_error("mask and image dimensions differ");
    std::vector<unsigned char> mask(g.voxels());
    for (size_t i = 0; i < mask.size(); ++i) mask[i] = mask16->GetBufferPointer()[i] > 0 ? 1 : 0;  // Clamp(0,1)

    // ---- ROIs: DenseROIGenerator over the ROI mask, raster order (x fastest) ----
    std::vector<unsigned char> roiMask(mask);
    if (!roiMaskPath.empty()) {
      std::cout << "Using ROI mask." << std::endl;
      auto rm = ife::nifti::Read<unsigned short>(roiMaskPath);
      if (rm->GetSize() != image->GetSize()) throw std::runtime_error("ROI mask and image dimensions differ");
      for (size_t i = 0; i < roiMask.size(); ++i) roiMask[i] = rm->GetBufferPointer()[i] == roiMaskValue ? 1 : 0;
    }
    std::vector<ife::Region> rois;
    {
      const int sx = (int)roiSize[0], sy = (int)roiSize[1], sz = (int)roiSize[2];
      size_t i = 0;
      for (int z = 0; z < g.size[2]; ++z)
        for (int y = 0; y < g.size[1]; ++y)
          for (int x = 0; x < g.size[0]; ++x, ++i) {
            if (!roiMask[i]) continue;
            const ife::Region r{{x - sx / 2, y - sy / 2, z - sz / 2, sx, sy, sz}};
            if (r[0] < 0 || r[1] < 0 || r[2] < 0 || r[0] + sx > g.size[0] || r[1] + sy > g.size[1] || r[2] + sz > g.size[2]) continue;
            rois.push_back(r);
          }
      std::ofstream out(ife::Path::join(outDirPath, prefix + ".ROIInfo"));
      ife::ROIReader::write(out, rois);
      if (!out.good()) { std::cerr << "Error writing ROI info file" << std::endl; return EXIT_FAILURE; }
    }

    // ---- histogram specification: one line of comma-separated edges per (scale, feature) ----
    std::ifstream isHist(histPath);
    if (!isHist.good()) { std::cerr << "Could not read histogram file '" << histPath << "'" << std::endl; return EXIT_FAILURE; }
    std::vector<float> edges;
    size_t histSize = 0, nHist = 0;
    while (isHist.good()) {
      std::string line;
      std::getline(isHist, line);
      if (line.empty()) { std::cout << "Empty line. Breaking" << std::endl; break; }
      if (line[0] == '#') { std::cout << "Skipping a line" << std::endl; continue; }
      std::stringstream ss(line);
      std::vector<float> row;
      float e;
      while (ss >> e) { row.push_back(e); ss.ignore(std::numeric_limits<std::streamsize>::max(), ','); }
      if (histSize == 0) histSize = row.size() + 1;
      else if (histSize != row.size() + 1) {
        std::cerr << "Histograms must have the same bin count" << std::endl
                  << "Expected " << histSize << " Got " << row.size() << std::endl
                  << "Number of histograms " << nHist + 1 << std::endl;
        return EXIT_FAILURE;
      }
      edges.insert(edges.end(), row.begin(), row.end());
      ++nHist;
    }
    if (nHist != numFeatures * scales.size()) {
      std::cerr << "Number of histograms must match number of features times number of scales" << std::endl
                << "Number of histograms = " << nHist << std::endl
                << "Number of features*scales = " << numFeatures * scales.size() << std::endl;
      return EXIT_FAILURE;
    }
    const size_t totalBins = histSize * nHist;

    // ---- slices of ROIs: all scales and features of a slice in one GPU call ----
    for (double s : scales) std::cout << "Processing scale " << s << std::endl;
    std::ofstream out(ife::Path::join(outDirPath, prefix + ".bag"));
    ife::CudaContext& c = ife::CudaContext::Instance();
    const size_t slice = 32768;
    std::vector<int> roiFlat;
    std::vector<uint32_t> counts;
    for (size_t j0 = 0; j0 < rois.size(); j0 += slice) {
      const size_t nj = std::min(slice, rois.size() - j0);
      roiFlat.clear();
      for (size_t j = 0; j < nj; ++j) roiFlat.insert(roiFlat.end(), rois[j0 + j].begin(), rois[j0 + j].end());
      counts.assign(nj * totalBins, 0u);
      c.Check(ife_cuda_emphysema_histograms(c.Handle(), image->GetBufferPointer(), mask.data(), g.size.data(), g.spacing.data(),
                                            scales.data(), (int)scales.size(), edges.data(), (int)histSize - 1, roiFlat.data(),
                                            (int)nj, counts.data(), IFE_MEM_HOST));
      // bag rows: frequencies = count / sum exactly as DenseHistogram::getFrequencies
      for (size_t j = 0; j < nj; ++j) {
        for (size_t h = 0; h < nHist; ++h) {
          const uint32_t* cnt = counts.data() + j * totalBins + h * histSize;
          int isum = 0;
          for (size_t l = 0; l < histSize; ++l) isum += (int)cnt[l];
          const float sum = (float)isum;
          for (size_t l = 0; l < histSize; ++l) {
            out << (float)cnt[l] / sum;
            if (h * histSize + l + 1 < totalBins) out << ",";
          }
        }
        out << '\n';
      }
    }
    if (!out.good()) { std::cerr << "Error writing histogram to file" << std::endl; return EXIT_FAILURE; }
  } catch (std::exception& e) {
    std::cerr << "Failed to process." << std::endl
              << "Image: " << imagePath << std::endl
              << "Mask: " << maskPath << std::endl
              << "ExceptionObject: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
