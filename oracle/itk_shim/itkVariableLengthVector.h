// TEST INFRASTRUCTURE ONLY (oracle).  Minimal stand-in for itk::VariableLengthVector so
// that the reference's ITK-light headers (Symmetric3x3EigenvalueSolver.h,
// EigenvalueFeaturesFunctor.h) compile UNMODIFIED from /root/reference/include into
// oracle/_ref/.  ITK itself is not installed in this image.  Only the members those
// headers use are provided; storage is heap-allocated per object like the real class,
// so the CPU baseline keeps the reference's per-voxel allocation cost.
#ifndef ORACLE_ITK_SHIM_VARIABLE_LENGTH_VECTOR_H
#define ORACLE_ITK_SHIM_VARIABLE_LENGTH_VECTOR_H
#include <cassert>
#include <cstddef>
#include <algorithm>

namespace itk {
template <typename T>
class VariableLengthVector {
public:
  typedef T ValueType;
  typedef unsigned int ElementIdentifier;

  VariableLengthVector() : m_Data(nullptr), m_Size(0) {}
  explicit VariableLengthVector(unsigned int n) : m_Data(new T[n]), m_Size(n) {}
  VariableLengthVector(const T* src, unsigned int n) : m_Data(new T[n]), m_Size(n) {
    std::copy(src, src + n, m_Data);
  }
  VariableLengthVector(const VariableLengthVector& o) : m_Data(new T[o.m_Size]), m_Size(o.m_Size) {
    std::copy(o.m_Data, o.m_Data + o.m_Size, m_Data);
  }
  VariableLengthVector(VariableLengthVector&& o) noexcept : m_Data(o.m_Data), m_Size(o.m_Size) {
    o.m_Data = nullptr;
    o.m_Size = 0;
  }
  VariableLengthVector& operator=(VariableLengthVector o) {
    std::swap(m_Data, o.m_Data);
    std::swap(m_Size, o.m_Size);
    return *this;
  }
  ~VariableLengthVector() { delete[] m_Data; }

  T& operator[](unsigned int i) { return m_Data[i]; }
  const T& operator[](unsigned int i) const { return m_Data[i]; }
  unsigned int Size() const { return m_Size; }
  unsigned int GetSize() const { return m_Size; }
  void Fill(const T& v) { std::fill(m_Data, m_Data + m_Size, v); }
  const T* GetDataPointer() const { return m_Data; }

private:
  T* m_Data;
  unsigned int m_Size;
};
}  // namespace itk
#endif
