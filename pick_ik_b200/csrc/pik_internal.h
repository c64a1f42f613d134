// pik_internal.h -- hooks between the translation units of libpik_b200.so (not part of the C-ABI).
#pragma once

#include <stddef.h>
#include <stdint.h>

#include "../../include/pik.h"

int pik_internal_solver_device(const pik_solver* s);
int pik_internal_solver_num_variables(const pik_solver* s);
void* pik_internal_solver_stream(const pik_solver* s);
// pik_solve_batch_async whose results stay in the solver's device buffers; pik_internal_finish = pik_solver_wait
int pik_internal_finish(pik_solver* s);
int pik_internal_solve_keep(pik_solver* s, const pik_params* params, int64_t B, int64_t first_problem_index,
                            const double* goal_pose, const double* seed, int64_t seed_stride, int32_t memory);
// packs those results into [B][n + 3] doubles (device, solver-owned; B < 0: rows of NaN for a failed shard);
// gather_elems > 0 also reserves a device block of that many doubles for the gathered result
int pik_internal_pack(pik_solver* s, int64_t B, size_t packed_elems, size_t gather_elems, double** packed, double** gather);
