// ROI text files as the reference reads and writes them (include/ife/IO/ROIReader.hxx:24-48,
// tools/MakeBag.cxx:285-291): one region per line, "[x, y, z][sx, sy, sz]", optional header
// line.  A region is {x0, y0, z0, sx, sy, sz}.
#ifndef IFE_B200_ROI_READER_H
#define IFE_B200_ROI_READER_H
#include <array>
#include <fstream>
#include <istream>
#include <limits>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace ife {

typedef std::array<int, 6> Region;

struct ROIReader {
  static std::vector<Region> read(std::istream& is, bool header) {
    const std::streamsize count = std::numeric_limits<std::streamsize>::max();
    std::vector<Region> rois;
    if (header) is.ignore(count, '\n');
    while (is.good()) {
      Region r;
      is.ignore(count, '[');
      is >> r[0]; is.ignore(count, ',');
      is >> r[1]; is.ignore(count, ',');
      is >> r[2]; is.ignore(count, '[');
      is >> r[3]; is.ignore(count, ',');
      is >> r[4]; is.ignore(count, ',');
      is >> r[5]; is.ignore(count, '\n');
      if (is.good()) rois.push_back(r);
    }
    return rois;
  }
  static std::vector<Region> read(const std::string& path, bool header) {
    std::ifstream is(path);
    if (!is.good()) throw std::runtime_error("cannot open ROI file '" + path + "'");
    return read(is, header);
  }
  static void write(std::ostream& os, const std::vector<Region>& rois) {
    for (const Region& r : rois)
      os << "[" << r[0] << ", " << r[1] << ", " << r[2] << "][" << r[3] << ", " << r[4] << ", " << r[5] << "]\n";
  }
};

}  // namespace ife
#endif
