// instaslice_host.hpp — C++ host-side mirror of the reference's allocator interface over the C ABI.
//
// The reference controller is Go and no Go toolchain exists in this image, so the host side above
// include/islplace.h is written in C++ with the reference's names, argument meaning and error behaviour
// (internal/controller/instaslice_controller.go at b34e86d):
//   InstasliceReconciler::findDeviceForASlice            :240-262
//   InstasliceReconciler::getStartIndexFromPreparedState :303-384  (occupancy build :306-328 = occupancyByte)
//   InstasliceReconciler::extractGpuProfile              :283-300
//   AllocationPolicy / FirstFitPolicy / LeftToRightPolicy / RightToLeftPolicy  :48-56, :436-469
//   InstasliceReconciler::PlacePending                   node loop + Prepared veto of Reconcile :188-232, batched
// Types follow api/v1alpha1/instaslice_types.go:23-72.  Nothing here decides a placement: every decision comes
// back from libislplace.so.  The Go twin of this file is integration/go/placement_engine.go.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "../../include/islplace.h"

namespace instaslice {

struct Placement { int Size = 0; int Start = 0; };
struct Mig {
    std::vector<Placement> Placements;
    std::string Profile;
    int Giprofileid = 0, CIProfileID = 0, CIEngProfileID = 0;
};
struct AllocationDetails {
    std::string Profile;
    uint32_t Start = 0, Size = 0;
    std::string PodUUID, GPUUUID, Nodename, Allocationstatus;
    int Giprofileid = 0, CIProfileID = 0, CIEngProfileID = 0;
    std::string Namespace, PodName;
};
struct PreparedDetails {
    std::string Profile;
    uint32_t Start = 0, Size = 0;
    std::string Parent, PodUUID;
    uint32_t Giinfoid = 0, Ciinfoid = 0;
};
struct InstasliceSpec {
    std::map<std::string, std::string> MigGPUUUID;
    std::map<std::string, AllocationDetails> Allocations;
    std::map<std::string, PreparedDetails> Prepared;
    std::vector<Mig> Migplacement;
};
struct Instaslice { std::string Name; InstasliceSpec Spec; };
struct InstasliceList { std::vector<Instaslice> Items; };
struct Pod { std::string UID, Namespace, Name; };

// :48-50 — the allocation-policy hook: packs, chooses nothing
struct AllocationPolicy {
    virtual ~AllocationPolicy() = default;
    virtual AllocationDetails SetAllocationDetails(const std::string& profileName, uint32_t newStart, uint32_t size, const std::string& podUUID,
                                                   const std::string& nodename, const std::string& processed, int discoveredGiprofile,
                                                   int Ciprofileid, int Ciengprofileid, const std::string& ns, const std::string& podName,
                                                   const std::string& gpuUuid) = 0;
};
struct FirstFitPolicy : AllocationPolicy {       // :436-453
    AllocationDetails SetAllocationDetails(const std::string&, uint32_t, uint32_t, const std::string&, const std::string&, const std::string&, int, int,
                                           int, const std::string&, const std::string&, const std::string&) override;
};
struct LeftToRightPolicy : AllocationPolicy {    // :456-461, a stub in the reference: empty AllocationDetails
    AllocationDetails SetAllocationDetails(const std::string&, uint32_t, uint32_t, const std::string&, const std::string&, const std::string&, int, int,
                                           int, const std::string&, const std::string&, const std::string&) override { return {}; }
};
struct RightToLeftPolicy : LeftToRightPolicy {}; // :464-469

enum class Verdict { Placed, None, Veto };       // allocation written / "failed to find allocatable gpu" everywhere / :198-203 requeue
struct PendingPod { Pod pod; std::string ProfileName; };
struct Outcome { Verdict verdict = Verdict::None; AllocationDetails alloc; };

extern const char* const kErrNoGpu;              // "failed to find allocatable gpu" (:261)

class InstasliceReconciler {
public:
    explicit InstasliceReconciler(uint32_t quirks = ISL_QUIRKS_REF_EXACT, uint32_t max_gpus = 1u << 16, uint32_t max_batch = 1u << 16);
    ~InstasliceReconciler();
    InstasliceReconciler(const InstasliceReconciler&) = delete;

    // Rebuild the flat inventory from the listed custom resources (the CR is the checkpoint).  Throws std::runtime_error
    // on what makes the reference panic and on engine errors.
    void Sync(const InstasliceList& list);
    // Incremental sync after ONE Instaslice object (list.Items[node]) changed: only that node's occupancy bytes are rewritten.
    // Falls back to Sync when the node's GPU set changed.
    void UpdateNode(const InstasliceList& list, size_t node);

    static uint8_t occupancyByte(const Instaslice& is, const std::string& gpuUUID);                        // :306-328
    uint32_t getStartIndexFromPreparedState(const Instaslice& is, const std::string& gpuUUID, const std::string& profileName);   // :303-384
    static void extractGpuProfile(const Instaslice& is, const std::string& profileName, int* size, int* gi, int* ci, int* cieng);  // :283-300
    // :240-262 — first GPU of ONE node; returns false and sets *err = kErrNoGpu when nothing fits.  Like the reference it
    // does not record the allocation (the tentative engine commit is released again).
    bool findDeviceForASlice(const InstasliceList& list, size_t node, const std::string& profileName, AllocationPolicy& policy, const Pod& pod,
                             AllocationDetails* out, std::string* err);
    // Reconcile's node loop for many gated pods in order, ONE engine call; allocations are written into `list`.
    std::vector<Outcome> PlacePending(InstasliceList& list, AllocationPolicy& policy, const std::vector<PendingPod>& pods);
    // The daemonset removed Allocations[podUID] (instaslice_daemonset.go:261-263).
    bool Release(InstasliceList& list, const std::string& podUID);

private:
    isl_engine* h_ = nullptr;
    std::vector<std::string> gpuUUID_;
    std::vector<size_t> gpuNode_;
    std::vector<uint32_t> nodeOff_;
    std::vector<uint8_t> nodeTable_;      // per-node profile table (heterogeneous clusters)
    uint32_t nTables_ = 1;
    std::map<std::string, uint8_t> profiles_;
    std::map<std::string, uint32_t> gpuIndex_;
    bool orphans_ = false;
    std::vector<isl_result> place(const std::vector<std::string>& names, uint32_t lo, uint32_t hi);
    void releaseSpan(const isl_result& r);
    Outcome commitOrVeto(InstasliceList& list, AllocationPolicy& policy, const PendingPod& p, const isl_result& r);
};

}  // namespace instaslice
