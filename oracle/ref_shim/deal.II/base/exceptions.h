// Minimal stand-in for <deal.II/base/exceptions.h>: Assert is a debug-mode check in deal.II (compiled out in release).
#pragma once
#include <stdexcept>
#include <string>
namespace dealii {
inline std::runtime_error ExcMessage(const std::string& s) { return std::runtime_error(s); }
inline std::runtime_error ExcNotImplemented() { return std::runtime_error("not implemented"); }
}  // namespace dealii
#define Assert(cond, exc) do { } while (0)
#define AssertThrow(cond, exc) do { if (!(cond)) throw (exc); } while (0)
