#pragma once
namespace dealii { namespace Utilities {
inline unsigned pow(unsigned base, int e) { unsigned r = 1; while (e-- > 0) r *= base; return r; }
}}  // namespace dealii::Utilities
