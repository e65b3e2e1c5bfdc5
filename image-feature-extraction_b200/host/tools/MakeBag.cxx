// MakeBag -i image -m mask -H histogram-spec -o outdir -s scale ... [-r roi-file] [-R bool]
//         [-M roi-mask] [-v roi-mask-value] [-n num-rois] [-x/-y/-z roi-size] [-p prefix] [-S seed]
// Flags, file formats and semantics of the reference tool (tools/MakeBag.cxx): for every scale
// the 8 features of ImageToEmphysemaFeaturesFilter over the whole volume, then per ROI every
// in-mask voxel's features are inserted into 8 DenseHistograms (one edge row per scale and
// feature, :327-371,448-457) and the frequencies (counts / sum, :462-470) become one CSV row
// per ROI in <outdir>/<prefix>.bag (:475-490).  ROIs come from a ROI file or are sampled at
// random around in-mask voxels (RegionOfInterestGenerator.hxx:20-59) and written to
// <outdir>/<prefix>.ROIInfo.  On the GPU the feature volumes are never materialised: one
// call bins all scales, features and ROIs (ife_cuda_emphysema_histograms).
// Addition: -S/--seed makes the random ROI sampling reproducible (the reference always
// reseeds from the clock).
#include <cstdlib>
#include <chrono>
#include <fstream>
#include <iostream>
#include <random>
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
  cmd.add("r", "roi-file", "Path to ROI file. If given the ROIs in this file will be used, otherwise ROIs will be generated.", false, "", "path");
  cmd.add("R", "roi-file-has-header", "Flag indicating if the ROI file has a header", false, "1", "boolean");
  cmd.add("M", "roi-mask", "Path to ROI mask file controlling the ROI generation (default: the image mask).", false, "", "path");
  cmd.add("v", "roi-mask-value", "Value in the ROI mask that should be used for inclusion.", false, "1", "MaskPixelType");
  cmd.add("n", "num-rois", "Number of ROIs to sample", false, "50", "N>=2");
  cmd.add("x", "roi-size-x", "Size of ROI in x dimension", false, "41", "N>=1");
  cmd.add("y", "roi-size-y", "Size of ROI in y dimension", false, "41", "N>=1");
  cmd.add("z", "roi-size-z", "Size of ROI in z dimension", false, "41", "N>=1");
  cmd.add("p", "prefix", "Prefix to use for output filenames", false, "", "string");
  cmd.add("S", "seed", "Seed for the random ROI sampling (default: non-deterministic)", false, "", "integer");
  int rc;
  if (!cmd.parse(argc, argv, &rc)) return rc;
  const std::string imagePath(cmd.value("image")), maskPath(cmd.value("mask")), histPath(cmd.value("histogram-spec"));
  const std::string outDirPath(cmd.value("outdir")), roiPath(cmd.value("roi-file")), roiMaskPath(cmd.value("roi-mask"));
  const std::string prefix(cmd.value("prefix"));
  std::vector<double> scales;
  for (const std::string& s : cmd.values("scale")) {
    float v;
    if (!ife::CmdLine::convert(s, &v)) { cmd.error("Couldn't read argument value from string '" + s + "'", "-s", &rc); return rc; }
    scales.push_back(v);
  }
  bool roiHasHeader = true;
  size_t numROIs = 50, roiSize[3] = {41, 41, 41};
  unsigned roiMaskValue = 1;
  if (!ife::CmdLine::to_bool(cmd.value("roi-file-has-header"), &roiHasHeader) ||
      !ife::CmdLine::convert(cmd.value("num-rois"), &numROIs) || !ife::CmdLine::convert(cmd.value("roi-size-x"), &roiSize[0]) ||
      !ife::CmdLine::convert(cmd.value("roi-size-y"), &roiSize[1]) || !ife::CmdLine::convert(cmd.value("roi-size-z"), &roiSize[2]) ||
      !ife::CmdLine::convert(cmd.value("roi-mask-value"), &roiMaskValue)) {
    cmd.error("Couldn't read a numeric argument value", "", &rc);
    return rc;
  }

  const auto t_start = std::chrono::steady_clock::now();
  auto since_start = [&t_start]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
  try {
    auto both = ife::nifti::ReadPair<float, unsigned short>(imagePath, maskPath);   // the two files are inflated concurrently
    const double t_read = since_start();
    ife::Image<float>::Pointer image = both.first;
    ife::Image<unsigned short>::Pointer mask16 = both.second;
    const ife::Geometry& g = image->GetGeometry();
    if (mask16->GetSize() != image->GetSize()) throw std::runtime_error("mask and image dimensions differ");
    std::vector<unsigned char> mask(g.voxels());
    for (size_t i = 0; i < mask.size(); ++i) mask[i] = mask16->GetBufferPointer()[i] > 0 ? 1 : 0;  // Clamp(0,1)

    // ---- ROIs ----
    std::vector<ife::Region> rois;
    if (roiPath.empty()) {
      std::vector<unsigned char> roiMask(mask);
      if (!roiMaskPath.empty()) {
        std::cout << "Using ROI mask." << std::endl;
        auto rm = ife::nifti::Read<unsigned short>(roiMaskPath);
        if (rm->GetSize() != image->GetSize()) throw std::runtime_error("ROI mask and image dimensions differ");
        for (size_t i = 0; i < roiMask.size(); ++i) roiMask[i] = rm->GetBufferPointer()[i] == roiMaskValue ? 1 : 0;
      }
      bool any = false;
      for (int z = (int)roiSize[2] / 2; z + (int)(roiSize[2] - roiSize[2] / 2) <= g.size[2] && !any; ++z)
        for (int y = (int)roiSize[1] / 2; y + (int)(roiSize[1] - roiSize[1] / 2) <= g.size[1] && !any; ++y)
          for (int x = (int)roiSize[0] / 2; x + (int)(roiSize[0] - roiSize[0] / 2) <= g.size[0]; ++x)
            if (roiMask[(size_t)x + (size_t)g.size[0] * (y + (size_t)g.size[1] * z)]) { any = true; break; }
      if (!any) throw std::runtime_error("Failed to generate ROIs: no in-mask voxel admits a ROI of this size inside the image");
      std::mt19937_64 gen;
      if (cmd.value("seed").empty()) gen.seed(std::random_device{}());
      else { unsigned long long s = 0; ife::CmdLine::convert(cmd.value("seed"), &s); gen.seed(s); }
      std::uniform_int_distribution<size_t> pick(0, g.voxels() - 1);
      while (rois.size() < numROIs) {
        const size_t i = pick(gen);
        if (!roiMask[i]) continue;
        const int x = (int)(i % g.size[0]), y = (int)((i / g.size[0]) % g.size[1]), z = (int)(i / ((size_t)g.size[0] * g.size[1]));
        const ife::Region r{{x - (int)(roiSize[0] / 2), y - (int)(roiSize[1] / 2), z - (int)(roiSize[2] / 2),
                             (int)roiSize[0], (int)roiSize[1], (int)roiSize[2]}};
        if (r[0] < 0 || r[1] < 0 || r[2] < 0 || r[0] + r[3] > g.size[0] || r[1] + r[4] > g.size[1] || r[2] + r[5] > g.size[2]) continue;
        rois.push_back(r);
      }
      std::ofstream out(ife::Path::join(outDirPath, prefix + ".ROIInfo"));
      ife::ROIReader::write(out, rois);
      if (!out.good()) { std::cerr << "Error writing ROI info file" << std::endl; return EXIT_FAILURE; }
    } else {
      rois = ife::ROIReader::read(roiPath, roiHasHeader);
      std::cout << "Got " << rois.size() << " rois." << std::endl;
    }
    if (rois.empty()) throw std::runtime_error("no ROIs");

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

    // ---- all scales, features and ROIs in one GPU call ----
    for (double s : scales) std::cout << "Processing scale " << s << std::endl;
    std::vector<int> roiFlat;
    for (const ife::Region& r : rois) roiFlat.insert(roiFlat.end(), r.begin(), r.end());
    std::vector<uint32_t> counts(rois.size() * totalBins);
    ife::CudaContext& c = ife::CudaContext::Instance();
    const auto t_gpu = std::chrono::steady_clock::now();
    c.Check(ife_cuda_emphysema_histograms(c.Handle(), image->GetBufferPointer(), mask.data(), g.size.data(), g.spacing.data(),
                                          scales.data(), (int)scales.size(), edges.data(), (int)histSize - 1, roiFlat.data(),
                                          (int)rois.size(), counts.data(), IFE_MEM_HOST));
    if (std::getenv("IFE_TIMING"))
      std::cerr << "[timing] inputs read at " << t_read << " s, GPU call done at " << since_start() << " s after start; reading (both files, summed): "
                << ife::nifti::read_seconds() << " s; page-locked allocation: " << ife::alloc_seconds()
                << " s; ife_cuda_emphysema_histograms(IFE_MEM_HOST, " << scales.size() << " scales, " << rois.size() << " ROIs): "
                << std::chrono::duration<double>(std::chrono::steady_clock::now() - t_gpu).count() << " s" << std::endl;

    // ---- bag: frequencies = count / sum exactly as DenseHistogram::getFrequencies ----
    std::ofstream out(ife::Path::join(outDirPath, prefix + ".bag"));
    for (size_t j = 0; j < rois.size(); ++j) {
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
    if (!out.good()) { std::cerr << "Error writing histogram to file" << std::endl; return EXIT_FAILURE; }
    if (std::getenv("IFE_TIMING")) std::cerr << "[timing] bag written at " << since_start() << " s after start" << std::endl;
  } catch (std::exception& e) {
    std::cerr << "Failed to process." << std::endl
              << "Image: " << imagePath << std::endl
              << "Mask: " << maskPath << std::endl
              << "ExceptionObject: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  if (std::getenv("IFE_TIMING")) std::cerr << "[timing] images released, main() returns at " << since_start() << " s after start" << std::endl;
  return EXIT_SUCCESS;
}
