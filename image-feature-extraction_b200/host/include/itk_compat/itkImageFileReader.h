// itk::ImageFileReader over the NIfTI-1 reader of ife/IO/NiftiIO.h: SetFileName / GetOutput /
// Update, with ITK's pull semantics (the file is read when a consumer updates, once per name).
#ifndef IFE_B200_ITK_COMPAT_IMAGE_FILE_READER_H
#define IFE_B200_ITK_COMPAT_IMAGE_FILE_READER_H
#include <memory>
#include <string>

#include "ife/IO/NiftiIO.h"
#include "itkImage.h"

namespace itk {
template <typename TOutputImage>
class ImageFileReader {
public:
  typedef ImageFileReader Self;
  typedef std::shared_ptr<Self> Pointer;
  typedef TOutputImage OutputImageType;
  static Pointer New() { return Pointer(new Self()); }
  void SetFileName(const std::string& name) { if (name != m_FileName) { m_FileName = name; m_Read = false; } }
  const std::string& GetFileName() const { return m_FileName; }
  OutputImageType* GetOutput() { return m_Output.get(); }
  void Update() {
    if (m_Read) return;
    try {
      auto img = ife::nifti::Read<typename OutputImageType::PixelType>(m_FileName);
      m_Output->SetGeometry(img->GetGeometry());
      m_Output->SetView(img->GetGeometry(), img->GetBufferPointer(), img->GetNumberOfPixels(), img);
    } catch (const std::exception& e) {
      throw ife::ExceptionObject(IFE_E_INVALID, e.what());
    }
    m_Read = true;
  }
private:
  ImageFileReader() : m_Output(OutputImageType::New()) { m_Output->SetSource([this]() { this->Update(); }); }
  std::string m_FileName;
  bool m_Read = false;
  typename OutputImageType::Pointer m_Output;
};
}  // namespace itk
#endif
