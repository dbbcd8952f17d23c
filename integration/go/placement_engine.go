// placement_engine.go — the cgo shim a maintainer adds to internal/controller/ of project-codeflare/instaslice
// (commit b34e86d) to route the allocator through libislplace.so.  SOURCE ONLY: there is no Go toolchain in the
// build image of this repository, so this file has never been compiled here; it is written against
// include/islplace.h and mirrors instaslice_b200/controller.py, which IS exercised by the GPU tests.
//
// What stays exactly as it is: Reconcile's pod state machine, the CRD types, the daemonset, and the
// AllocationPolicy hook (policy.SetAllocationDetails still packs the AllocationDetails).  What changes:
// the body of findDeviceForASlice / getStartIndexFromPreparedState (instaslice_controller.go:240-262,
// 303-384) becomes one isl_place_batch call.
package controller

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -lislplace -lcudart
#include <stdlib.h>
#include <cuda_runtime_api.h>
#include "islplace.h"
static void* isl_pinned(size_t n) { void* p = 0; return cudaHostAlloc(&p, n, cudaHostAllocDefault) == cudaSuccess ? p : 0; }
static void  isl_unpin(void* p)   { cudaFreeHost(p); }
*/
import "C"

import (
	"fmt"
	"reflect"
	"sort"
	"unsafe"

	inferencev1alpha1 "codeflare.dev/instaslice/api/v1alpha1"
	v1 "k8s.io/api/core/v1"
)

// PlacementEngine mirrors the listed Instaslice objects into the device-resident inventory.
type PlacementEngine struct {
	h        *C.isl_engine
	gpuUUID  []string          // canonical GPU index -> UUID
	gpuNode  []int             // canonical GPU index -> index into the Instaslice list
	profiles map[string]uint8  // profile name -> row index (FIRST Migplacement row with that name, :332-340)
	orphans  bool              // a realised slice outlived its allocation: the :198-203 veto can fire
	nodeOff  []C.uint32_t      // node index -> first canonical GPU index
}

func NewPlacementEngine(maxGPUs, maxBatch uint32) (*PlacementEngine, error) {
	cfg := C.isl_config{abi_version: C.ISL_ABI_VERSION, policy: C.ISL_POLICY_FIRST_FIT, quirks: C.ISL_QUIRKS_REF_EXACT,
		device: -1, max_gpus: C.uint32_t(maxGPUs), max_batch: C.uint32_t(maxBatch)}
	var h *C.isl_engine
	if rc := C.isl_create(&cfg, &h); rc != C.ISL_OK {
		return nil, fmt.Errorf("isl_create: %s", C.GoString(C.isl_strerror(rc)))
	}
	return &PlacementEngine{h: h}, nil
}

func (e *PlacementEngine) Close() { C.isl_destroy(e.h) }

// occupancyByte is instaslice_controller.go:306-328: dangling Prepared entries and every Allocations entry
// (any status) mark their slices.
func occupancyByte(is *inferencev1alpha1.Instaslice, gpuUUID string) (uint8, error) {
	var busy uint8
	for _, item := range is.Spec.Prepared {
		if item.Parent == gpuUUID && item.PodUUID == "" {
			if item.Start+item.Size > 8 {
				return 0, fmt.Errorf("prepared span beyond slice 7")
			}
			busy |= uint8(((1 << item.Size) - 1) << item.Start)
		}
	}
	for _, item := range is.Spec.Allocations {
		if item.GPUUUID == gpuUUID {
			if item.Start+item.Size > 8 {
				return 0, fmt.Errorf("allocation span beyond slice 7")
			}
			busy |= uint8(((1 << item.Size) - 1) << item.Start)
		}
	}
	return busy, nil
}

// Sync rebuilds the flat inventory from the custom resources (the CR is the checkpoint).  Canonical order:
// nodes in list order, GPUs by ascending UUID inside a node (the reference's orders are random, :85, :242).
func (e *PlacementEngine) Sync(list *inferencev1alpha1.InstasliceList) error {
	if len(list.Items) == 0 {
		return fmt.Errorf("no Instaslice objects")
	}
	// every node publishes its OWN Migplacement (instaslice_daemonset.go:588-664): group identical tables (<= 8), profile NAME
	// index = order of first appearance; rows[t*P + name] with n_starts == 0 when table t has no row of that name
	type tableT = []inferencev1alpha1.Mig
	tables := []tableT{}
	nodeTable := make([]C.uint8_t, len(list.Items))
	e.profiles = map[string]uint8{}
	for n := range list.Items {
		mig := list.Items[n].Spec.Migplacement
		t := 0
		for ; t < len(tables); t++ {
			if reflect.DeepEqual(tables[t], tableT(mig)) {
				break
			}
		}
		if t == len(tables) {
			if len(tables) >= int(C.ISL_MAX_TABLES) {
				return fmt.Errorf("more than %d distinct per-node profile tables", int(C.ISL_MAX_TABLES))
			}
			tables = append(tables, mig)
			for _, m := range mig {
				if len(m.Placements) == 0 {
					return fmt.Errorf("profile %s has no placements", m.Profile) // the reference panics at :334
				}
				if _, ok := e.profiles[m.Profile]; !ok {
					e.profiles[m.Profile] = uint8(len(e.profiles))
				}
			}
		}
		nodeTable[n] = C.uint8_t(t)
	}
	P := len(e.profiles)
	rows := make([]C.isl_profile, len(tables)*P)
	for t, mig := range tables {
		seenName := map[string]bool{}
		for _, m := range mig {
			if seenName[m.Profile] { // the start search uses the FIRST row with a name (:332-340)
				continue
			}
			seenName[m.Profile] = true
			r := &rows[t*P+int(e.profiles[m.Profile])]
			r.size = C.uint8_t(m.Placements[0].Size)
			seen := map[int]bool{}
			for _, p := range m.Placements {
				if !seen[p.Start] {
					seen[p.Start] = true
					r.starts[r.n_starts] = C.uint8_t(p.Start)
					r.n_starts++
				}
			}
			r.gi_profile_id, r.ci_profile_id, r.ci_eng_profile_id = C.int32_t(m.Giprofileid), C.int32_t(m.CIProfileID), C.int32_t(m.CIEngProfileID)
		}
	}
	nodeOff := []C.uint32_t{0}
	occ := []C.uint8_t{}
	e.gpuUUID, e.gpuNode, e.orphans = nil, nil, false
	for n := range list.Items {
		is := &list.Items[n]
		uuids := make([]string, 0, len(is.Spec.MigGPUUUID))
		for u := range is.Spec.MigGPUUUID {
			uuids = append(uuids, u)
		}
		sort.Strings(uuids)
		for _, u := range uuids {
			b, err := occupancyByte(is, u)
			if err != nil {
				return err
			}
			occ = append(occ, C.uint8_t(b))
			e.gpuUUID = append(e.gpuUUID, u)
			e.gpuNode = append(e.gpuNode, n)
		}
		nodeOff = append(nodeOff, C.uint32_t(len(e.gpuUUID)))
		for _, p := range is.Spec.Prepared {
			if _, live := is.Spec.Allocations[p.PodUUID]; p.PodUUID != "" && !live {
				e.orphans = true
			}
		}
	}
	if rc := C.isl_load_profile_tables(e.h, C.uint32_t(len(tables)), C.uint32_t(P), &rows[0]); rc != C.ISL_OK {
		return fmt.Errorf("isl_load_profile_tables: %s", C.GoString(C.isl_strerror(rc)))
	}
	if rc := C.isl_load_inventory(e.h, C.uint32_t(len(list.Items)), &nodeOff[0], &occ[0]); rc != C.ISL_OK {
		return fmt.Errorf("isl_load_inventory: %s", C.GoString(C.isl_strerror(rc)))
	}
	if rc := C.isl_set_node_tables(e.h, C.uint32_t(len(list.Items)), &nodeTable[0]); rc != C.ISL_OK {
		return fmt.Errorf("isl_set_node_tables: %s", C.GoString(C.isl_strerror(rc)))
	}
	e.nodeOff = nodeOff
	return nil
}

// UpdateNode is the incremental sync after ONE Instaslice object changed (an Allocations / Prepared entry appeared or was
// deleted): only that node's occupancy bytes are rewritten (isl_write_occupancy) instead of re-listing the cluster (:85).
func (e *PlacementEngine) UpdateNode(list *inferencev1alpha1.InstasliceList, n int) error {
	is := &list.Items[n]
	lo, hi := int(e.nodeOff[n]), int(e.nodeOff[n+1])
	if len(is.Spec.MigGPUUUID) != hi-lo {
		return e.Sync(list)
	}
	occ := make([]C.uint8_t, 0, hi-lo)
	for g := lo; g < hi; g++ {
		if _, ok := is.Spec.MigGPUUUID[e.gpuUUID[g]]; !ok {
			return e.Sync(list)
		}
		b, err := occupancyByte(is, e.gpuUUID[g])
		if err != nil {
			return err
		}
		occ = append(occ, C.uint8_t(b))
	}
	if rc := C.isl_write_occupancy(e.h, C.uint32_t(lo), C.uint32_t(len(occ)), &occ[0]); rc != C.ISL_OK {
		return fmt.Errorf("isl_write_occupancy: %s", C.GoString(C.isl_strerror(rc)))
	}
	return nil
}

// PendingPod is one gated pod that has no allocation yet (Reconcile :148-187 already ran for it).
type PendingPod struct {
	Pod         *v1.Pod
	ProfileName string // r.extractProfileName(limits), :154
}

// PlacePending resolves the pods in order with ONE engine call and packs the answers with the unchanged policy
// hook.  result[i] == nil means "failed to find allocatable gpu" on every node (:261, :229-232: requeue).
func (r *InstasliceReconciler) PlacePending(e *PlacementEngine, list *inferencev1alpha1.InstasliceList, policy AllocationPolicy,
	pods []PendingPod) ([]*inferencev1alpha1.AllocationDetails, error) {
	n := len(pods)
	out := make([]*inferencev1alpha1.AllocationDetails, n)
	if n == 0 {
		return out, nil
	}
	if e.orphans && n > 1 { // the exact-match veto (:198-203) must see one pod at a time
		for i := range pods {
			one, err := r.PlacePending(e, list, policy, pods[i:i+1])
			if err != nil {
				return nil, err
			}
			out[i] = one[0]
		}
		return out, nil
	}
	// C-allocated request/result arrays: no Go pointer is retained by the engine after the call returns
	req := (*[1 << 28]C.isl_request)(C.malloc(C.size_t(n) * C.sizeof_isl_request))[:n:n]
	res := (*[1 << 28]C.isl_result)(C.malloc(C.size_t(n) * C.sizeof_isl_result))[:n:n]
	defer C.free(unsafe.Pointer(&req[0]))
	defer C.free(unsafe.Pointer(&res[0]))
	for i, p := range pods {
		row, ok := e.profiles[p.ProfileName]
		if !ok {
			row = C.ISL_PROFILE_UNKNOWN
		}
		req[i] = C.isl_request{handle: C.uint32_t(i), profile: C.uint8_t(row), op: C.ISL_OP_ALLOC}
	}
	if rc := C.isl_place_batch(e.h, C.uint32_t(n), &req[0], &res[0]); rc != C.ISL_OK {
		return nil, fmt.Errorf("isl_place_batch: %s (%s)", C.GoString(C.isl_strerror(rc)), C.GoString(C.isl_last_cuda_error(e.h)))
	}
	for i, p := range pods {
		if res[i].status != C.ISL_ST_PLACED {
			continue
		}
		gpu := int(res[i].gpu)
		is := &list.Items[e.gpuNode[gpu]]
		size, gi, ci, cieng := r.extractGpuProfile(is, p.ProfileName) // :283-300, unchanged
		a := policy.SetAllocationDetails(p.ProfileName, uint32(res[i].start), uint32(size), string(p.Pod.UID), is.Name, "creating",
			gi, ci, cieng, p.Pod.Namespace, p.Pod.Name, e.gpuUUID[gpu]) // :254-256, unchanged
		vetoed := false
		for _, item := range is.Spec.Prepared { // :198-203
			if item.Parent == a.GPUUUID && item.Size == a.Size && item.Start == a.Start {
				vetoed = true
			}
		}
		if vetoed {
			span := C.isl_span{gpu: res[i].gpu, start: res[i].start, size: res[i].size}
			C.isl_free_batch(e.h, 1, &span) // undo the tentative commit; the caller requeues after 1 s
			continue
		}
		out[i] = a
	}
	return out, nil
}

// Release tells the engine that the daemonset removed Allocations[podUID] (instaslice_daemonset.go:261-263).
func (e *PlacementEngine) Release(gpuIndex uint32, start, size uint8) {
	span := C.isl_span{gpu: C.uint32_t(gpuIndex), start: C.uint8_t(start), size: C.uint8_t(size)}
	C.isl_free_batch(e.h, 1, &span)
}

// PlaceBacklog resolves several ordered batches (e.g. per-namespace queues drained in turn) with ONE isl_place_stream
// call: identical answers to PlacePending batch after batch, pipelined on the device.  The request / result arrays are
// pinned (cudaHostAlloc through the tiny C helper below, or cudaHostRegister on C.malloc'ed arrays kept for the life of
// the controller): the engine then copies batch b while it already places batch b-1 and writes finished chunks straight
// into `res`.  Pageable arrays work as well, without that overlap.
//
//   // in the cgo preamble:
//   //   #include <cuda_runtime_api.h>
//   //   static void* isl_pinned(size_t n) { void* p = 0; return cudaHostAlloc(&p, n, cudaHostAllocDefault) == cudaSuccess ? p : 0; }
//   //   static void  isl_unpin(void* p)   { cudaFreeHost(p); }
func (r *InstasliceReconciler) PlaceBacklog(e *PlacementEngine, list *inferencev1alpha1.InstasliceList, policy AllocationPolicy,
	batches [][]PendingPod) ([][]*inferencev1alpha1.AllocationDetails, error) {
	total := 0
	sizes := make([]C.uint32_t, len(batches))
	for b, pods := range batches {
		sizes[b] = C.uint32_t(len(pods))
		total += len(pods)
	}
	out := make([][]*inferencev1alpha1.AllocationDetails, len(batches))
	if total == 0 || e.orphans { // the exact-match veto (:198-203) needs one pod at a time: fall back to PlacePending
		for b, pods := range batches {
			one, err := r.PlacePending(e, list, policy, pods)
			if err != nil {
				return nil, err
			}
			out[b] = one
		}
		return out, nil
	}
	req := (*[1 << 28]C.isl_request)(C.isl_pinned(C.size_t(total) * C.sizeof_isl_request))[:total:total]
	res := (*[1 << 28]C.isl_result)(C.isl_pinned(C.size_t(total) * C.sizeof_isl_result))[:total:total]
	defer C.isl_unpin(unsafe.Pointer(&req[0]))
	defer C.isl_unpin(unsafe.Pointer(&res[0]))
	i := 0
	for _, pods := range batches {
		for _, p := range pods {
			row, ok := e.profiles[p.ProfileName]
			if !ok {
				row = C.ISL_PROFILE_UNKNOWN
			}
			req[i] = C.isl_request{handle: C.uint32_t(i), profile: C.uint8_t(row), op: C.ISL_OP_ALLOC}
			i++
		}
	}
	if rc := C.isl_place_stream(e.h, C.uint32_t(len(batches)), &sizes[0], &req[0], &res[0]); rc != C.ISL_OK {
		return nil, fmt.Errorf("isl_place_stream: %s (%s)", C.GoString(C.isl_strerror(rc)), C.GoString(C.isl_last_cuda_error(e.h)))
	}
	i = 0
	for b, pods := range batches {
		out[b] = make([]*inferencev1alpha1.AllocationDetails, len(pods))
		for k, p := range pods {
			if res[i].status == C.ISL_ST_PLACED {
				gpu := int(res[i].gpu)
				is := &list.Items[e.gpuNode[gpu]]
				size, gi, ci, cieng := r.extractGpuProfile(is, p.ProfileName) // :283-300, unchanged
				out[b][k] = policy.SetAllocationDetails(p.ProfileName, uint32(res[i].start), uint32(size), string(p.Pod.UID), is.Name, "creating",
					gi, ci, cieng, p.Pod.Namespace, p.Pod.Name, e.gpuUUID[gpu]) // :254-256, unchanged
			}
			i++
		}
	}
	return out, nil
}

// WhatIf runs `plan` against a device-side snapshot of the occupancy and puts the snapshot back: defragmentation planning
// ("would these pods fit if those slices were released?") without touching the live state (isl_snapshot_occupancy /
// isl_restore_occupancy: a 1-byte-per-GPU device copy).
func (e *PlacementEngine) WhatIf(plan func() error) error {
	if rc := C.isl_snapshot_occupancy(e.h); rc != C.ISL_OK {
		return fmt.Errorf("isl_snapshot_occupancy: %s", C.GoString(C.isl_strerror(rc)))
	}
	err := plan()
	if rc := C.isl_restore_occupancy(e.h); rc != C.ISL_OK && err == nil {
		err = fmt.Errorf("isl_restore_occupancy: %s", C.GoString(C.isl_strerror(rc)))
	}
	return err
}
