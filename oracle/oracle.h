/*
 * oracle.h — C API of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker or as the timed
 * CPU baseline.  The product path (libislplace.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference (Go) cannot be built in this image (no go /
 * gccgo, un-vendored deps, no network) and none of its own tests reaches the
 * allocator.  What pins this oracle instead:
 *   (1) the hand-derived known-answer vectors of SURVEY.md 8c (tests/golden/),
 *   (2) three independent restatements that must agree on randomized inputs:
 *       ref_faithful.cpp (structure-for-structure), ref_fast.cpp (bitmask),
 *       ref_py.py (pure Python, dict-shaped CRD objects).
 *
 * The record layouts are shared with the product header on purpose: the oracle
 * answers the very same isl_request/isl_result arrays.
 */
#ifndef ISL_ORACLE_H
#define ISL_ORACLE_H

#include <stdint.h>
#include "../include/islplace.h"

#ifdef __cplusplus
extern "C" {
#endif

/* extra result status only the faithful driver can produce (Reconcile :198-203) */
#define ORC_ST_VETO_REQUEUE 100u

/* ---- single (occupancy byte, profile row) evaluation: :343-383 ---------- */
uint8_t orc_start_for(const isl_profile* row, uint32_t quirks, uint8_t occ);

/* ---- ref_faithful: string-keyed CRD objects, per-pod rescans ------------ */
typedef struct orc_faithful orc_faithful;
orc_faithful* orc_f_new(uint32_t n_nodes, const uint32_t* node_off,
                        uint32_t n_profiles, const isl_profile* rows, uint32_t quirks);
/* heterogeneous cluster: rows[t * n_profiles + p] (n_starts == 0: table t has no row of that name), node n uses table node_table[n] */
orc_faithful* orc_f_new_tables(uint32_t n_nodes, const uint32_t* node_off, uint32_t n_tables, uint32_t n_profiles, const isl_profile* rows,
                               const uint8_t* node_table, uint32_t quirks);
void     orc_f_delete(orc_faithful* h);
/* Spec.Prepared entry on canonical GPU `gpu`; pod_id < 0 => PodUUID == "" (dangling, counted as busy, :313) */
int      orc_f_add_prepared(orc_faithful* h, uint32_t gpu, uint32_t start, uint32_t size, int64_t pod_id);
/* Spec.Allocations[pod-<pod_id>] on canonical GPU `gpu` (any status counts, :322-328) */
int      orc_f_add_allocation(orc_faithful* h, uint32_t gpu, uint32_t start, uint32_t size, uint64_t pod_id);
/* Canonical batch: all FREEs, then ALLOCs in order (one Reconcile each).  all_nodes != 0 reproduces the
 * reference's missing `break` (Q5): the pod is allocated on every node with capacity; out[] reports the first. */
int      orc_f_place(orc_faithful* h, uint32_t n, const isl_request* in, isl_result* out, int all_nodes);
int      orc_f_place_batch(orc_faithful* h, uint32_t n, const isl_request* in, isl_result* out);   /* orc_f_place(..., all_nodes = 0) */
void     orc_f_occupancy(orc_faithful* h, uint8_t* out);      /* :306-328 for every GPU */
uint64_t orc_f_num_allocations(orc_faithful* h);

/* ---- ref_fast: 8-bit occupancy + first-start table + moving pointers ----- */
typedef struct orc_fast orc_fast;
orc_fast* orc_fast_new(uint32_t n_nodes, const uint32_t* node_off,
                       uint32_t n_profiles, const isl_profile* rows, uint32_t quirks, uint32_t policy);
orc_fast* orc_fast_new_tables(uint32_t n_nodes, const uint32_t* node_off, uint32_t n_tables, uint32_t n_profiles, const isl_profile* rows,
                              const uint8_t* node_table, uint32_t quirks, uint32_t policy);
void     orc_fast_delete(orc_fast* h);
void     orc_fast_load(orc_fast* h, const uint8_t* occ);
int      orc_fast_place(orc_fast* h, uint32_t n, const isl_request* in, isl_result* out);
void     orc_fast_occupancy(orc_fast* h, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
