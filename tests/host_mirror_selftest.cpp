// Self-test of the C++ host mirror (instaslice_b200/host) on a GPU: the hand-derived vectors of SURVEY.md 8c
// (instaslice_controller.go:240-384) and the README trace, through the reference-named interface.
// Built and run by tests/test_gpu_host_mirror.py.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../instaslice_b200/host/instaslice_host.hpp"

using namespace instaslice;

#define EXPECT(cond)                                                             \
    do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)

static std::vector<Mig> a100_40gb() {
    struct R { const char* n; int size; std::vector<int> starts; int gi; };
    const std::vector<R> rows = {{"1g.5gb", 1, {0, 1, 2, 3, 4, 5, 6}, 0}, {"2g.10gb", 2, {0, 2, 4}, 1}, {"3g.20gb", 4, {0, 4}, 2},
                                 {"4g.20gb", 4, {0}, 3},                  {"7g.40gb", 8, {0}, 4},        {"1g.10gb", 2, {0, 2, 4, 6}, 9}};
    std::vector<Mig> out;
    for (const R& r : rows) {
        Mig m; m.Profile = r.n; m.Giprofileid = r.gi; m.CIProfileID = r.gi; m.CIEngProfileID = 0;
        for (int s : r.starts) m.Placements.push_back({r.size, s});
        out.push_back(m);
    }
    return out;
}

static Instaslice node(const std::string& name, const std::vector<std::string>& gpus) {
    Instaslice is; is.Name = name; is.Spec.Migplacement = a100_40gb();
    for (const std::string& g : gpus) is.Spec.MigGPUUUID[g] = "NVIDIA A100-PCIE-40GB";
    return is;
}

int main() {
    FirstFitPolicy policy;
    {   // BASELINE config 1: samples/test-pod.yaml (nvidia.com/mig-1g.5gb) on one empty GPU
        InstasliceList list; list.Items.push_back(node("kind-control-plane", {"GPU-31cfe05c"}));
        InstasliceReconciler r; r.Sync(list);
        AllocationDetails a; std::string err;
        Pod pod{"uid-1", "default", "cuda-vectoradd-1"};
        EXPECT(r.findDeviceForASlice(list, 0, "1g.5gb", policy, pod, &a, &err));
        EXPECT(a.Profile == "1g.5gb" && a.Start == 0 && a.Size == 1 && a.Giprofileid == 0 && a.CIProfileID == 0 && a.CIEngProfileID == 0);
        EXPECT(a.Allocationstatus == "creating" && a.GPUUUID == "GPU-31cfe05c" && a.Nodename == "kind-control-plane" && a.PodUUID == "uid-1");
        EXPECT(!r.findDeviceForASlice(list, 0, "7g.40gb", policy, pod, &a, &err) && err == kErrNoGpu);     // Q1: 7g never places
        EXPECT(r.getStartIndexFromPreparedState(list.Items[0], "GPU-31cfe05c", "3g.20gb") == 0);
        EXPECT(r.getStartIndexFromPreparedState(list.Items[0], "GPU-31cfe05c", "9g.99gb") == 9);
        LeftToRightPolicy stub;
        EXPECT(stub.SetAllocationDetails("x", 1, 1, "", "", "", 0, 0, 0, "", "", "").Profile.empty());
    }
    {   // sequence vectors on one empty GPU
        InstasliceList list; list.Items.push_back(node("n0", {"GPU-0"}));
        InstasliceReconciler r; r.Sync(list);
        const std::vector<std::string> profs = {"3g.20gb", "1g.5gb", "2g.10gb", "3g.20gb", "1g.5gb", "1g.5gb"};
        const std::vector<int> want = {0, 4, 9, 9, 5, 6};
        std::vector<PendingPod> pods;
        for (size_t i = 0; i < profs.size(); ++i) pods.push_back({Pod{"u" + std::to_string(i), "default", "p" + std::to_string(i)}, profs[i]});
        std::vector<Outcome> out = r.PlacePending(list, policy, pods);
        for (size_t i = 0; i < out.size(); ++i) {
            if (want[i] == 9) EXPECT(out[i].verdict == Verdict::None);
            else EXPECT(out[i].verdict == Verdict::Placed && (int)out[i].alloc.Start == want[i] && out[i].alloc.GPUUUID == "GPU-0");
        }
        EXPECT(InstasliceReconciler::occupancyByte(list.Items[0], "GPU-0") == 0x7F);
        EXPECT(list.Items[0].Spec.Allocations.size() == 4);
        // release the 3g and place another one: frees reach the engine
        EXPECT(r.Release(list, "u0"));
        out = r.PlacePending(list, policy, {{Pod{"u9", "default", "p9"}, "4g.20gb"}});
        EXPECT(out[0].verdict == Verdict::Placed && out[0].alloc.Start == 0 && out[0].alloc.Giprofileid == 3);
    }
    {   // README.md:166-174,236-243: a dangling 3g.20gb on each of two GPUs, the 1g pod lands beside it on GPU 0;
        // second node only used when the first is full (canonical multi-node semantics: first node wins)
        InstasliceList list;
        list.Items.push_back(node("node-a", {"GPU-a0", "GPU-a1"}));
        list.Items.push_back(node("node-b", {"GPU-b0"}));
        for (const char* g : {"GPU-a0", "GPU-a1"}) {
            PreparedDetails p; p.Profile = "3g.20gb"; p.Start = 0; p.Size = 4; p.Parent = g;       // PodUUID empty: dangling
            list.Items[0].Spec.Prepared[std::string("MIG-") + g] = p;
        }
        InstasliceReconciler r; r.Sync(list);
        std::vector<PendingPod> pods;
        for (int i = 0; i < 8; ++i) pods.push_back({Pod{"v" + std::to_string(i), "default", "q"}, "1g.5gb"});
        const std::vector<Outcome> out = r.PlacePending(list, policy, pods);
        const char* gpus[] = {"GPU-a0", "GPU-a0", "GPU-a0", "GPU-a1", "GPU-a1", "GPU-a1", "GPU-b0", "GPU-b0"};
        const int starts[] = {4, 5, 6, 4, 5, 6, 0, 1};
        for (int i = 0; i < 8; ++i) EXPECT(out[i].verdict == Verdict::Placed && out[i].alloc.GPUUUID == gpus[i] && (int)out[i].alloc.Start == starts[i]);
        EXPECT(out[6].alloc.Nodename == "node-b" && list.Items[1].Spec.Allocations.size() == 2);
    }
    {   // the exact-match Prepared veto (:198-203): a realised slice whose allocation is gone blocks exactly its span
        InstasliceList list; list.Items.push_back(node("n0", {"GPU-0"}));
        PreparedDetails p; p.Profile = "1g.5gb"; p.Start = 0; p.Size = 1; p.Parent = "GPU-0"; p.PodUUID = "gone";
        list.Items[0].Spec.Prepared["MIG-x"] = p;
        InstasliceReconciler r; r.Sync(list);
        std::vector<Outcome> out = r.PlacePending(list, policy, {{Pod{"w0", "default", "q"}, "1g.5gb"}, {Pod{"w1", "default", "q"}, "2g.10gb"}});
        EXPECT(out[0].verdict == Verdict::Veto && out[1].verdict == Verdict::Placed && out[1].alloc.Start == 0);
    }
    printf("host mirror selftest: PASS\n");
    return 0;
}
