// Minimal stand-in for <deal.II/base/config.h>: just what the reference headers compiled by oracle/Makefile `ref` use.
// deal.II itself is not available in this image; see oracle/ref_shim/README.md.
#pragma once
#define DEAL_II_ALWAYS_INLINE __attribute__((always_inline))
