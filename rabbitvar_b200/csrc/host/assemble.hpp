// assemble.hpp — host-side string assembly of scored variants (stub, filled in below).
#pragma once
#include "pileup_model.hpp"
#include "../kernels/rv_core.cuh"
#include <stdio.h>

namespace rvhost {
inline void dump_variants(FILE* out, const rv_params& P, const std::vector<rv_variant>& v,
                          const std::vector<std::vector<rv_patch_entry> >& patches, const std::vector<rv_region>& regs,
                          const rvk::RefView& ref, const std::string& chr) {
  (void)out; (void)P; (void)v; (void)patches; (void)regs; (void)ref; (void)chr;
}
}  // namespace rvhost
