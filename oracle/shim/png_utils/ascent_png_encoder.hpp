/* Stub for src/libs/png_utils/ascent_png_encoder.hpp (the real header needs conduit.hpp, which
 * is not in this container).  apcomp::Image::Save/SaveDepth are the only users and the oracle
 * never calls them: images are compared as arrays from Python.  Test infrastructure only. */
#ifndef ORACLE_SHIM_PNG_ENCODER_HPP
#define ORACLE_SHIM_PNG_ENCODER_HPP
#include <string>
namespace ascent
{
class PNGEncoder
{
public:
  void Encode(const unsigned char*, int, int) {}
  void Encode(const float*, int, int) {}
  void Save(const std::string&) {}
};
}
#endif
