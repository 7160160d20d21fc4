// Host-side plan compiler interface (see plan_build.cpp).
#pragma once
#include <stddef.h>

#include <string>
#include <vector>

#include "plan.h"

// SerdeFormat of the reference (helpers.rs:7-19)
#define H2V_FMT_PROCESSED 0
#define H2V_FMT_RAW_BYTES 1
#define H2V_FMT_RAW_BYTES_UNCHECKED 2

namespace h2v {

struct PlanInfo {
  u32 k, n_points, n_scalars, n_challenges, proof_len, n_inst_cols, n_shared, n_mo;
};

int build_plan(const u8* params, size_t params_len, int params_fmt, const u8* vk, size_t vk_len, int vk_fmt, int multiopen,
               int hash, std::vector<u8>& blob, PlanInfo& info, std::string& err);

}  // namespace h2v
