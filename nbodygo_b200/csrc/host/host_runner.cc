// host_runner.cc — GpuStepper (the C++ twin of the cgo shim) and ComputationRunner.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <stdexcept>

#include "nbody_host.h"

namespace nbodygo {

// ---------------------------------------------------------------- GpuStepper
GpuStepper::GpuStepper(int device, int64_t capacity) : device_(device) { recreate(capacity, 0); }

void GpuStepper::recreate(int64_t capacity, int64_t pairCapacity)
{
    if (h_) { nb_destroy(h_); h_ = nullptr; }
    const int rc = nb_create(device_, capacity, pairCapacity, &h_);
    if (rc != NB_OK) {
        // no CPU fallback by design: the caller must not silently continue on the host
        throw std::runtime_error(std::string("[ERROR] nb_create: ") + nb_last_error(nullptr));
    }
    cap_ = capacity;
    pairCap_ = pairCapacity > 0 ? pairCapacity : 4 * capacity + 65536;  // nb_create's default
    n_ = 0;
    dirty_ = true;
    // the per-cycle Renderable snapshot lands in pinned host memory as part of the cycle itself
    if (nb_render_buffers(h_, &pinXyz_, &pinExists_) != NB_OK) { pinXyz_ = nullptr; pinExists_ = nullptr; }
}

GpuStepper::~GpuStepper()
{
    if (h_) nb_destroy(h_);
}

void GpuStepper::grow(size_t n)
{
    for (auto *v : {&x, &y, &z, &vx, &vy, &vz, &mass, &radius, &rest, &ff, &fs})
        if (v->size() < n) v->resize(n);
    if (beh.size() < n) { beh.resize(n); flags.resize(n); exists.resize(n); xyz.resize(3 * n); }
}

static uint8_t flagsOf(const Body &b)
{
    uint8_t f = 0;
    if (b.Exists) f |= NB_F_EXISTS;
    if (b.fragmenting) f |= NB_F_FRAGMENTING;
    if (b.Pinned) f |= NB_F_PINNED;
    if (b.IsSun) f |= NB_F_SUN;
    if (b.WithTelemetry) f |= NB_F_TELEMETRY;
    return f;
}

void GpuStepper::upload(BodyCollection &bc)
{
    auto &arr = bc.GetArray();
    const size_t n = arr.size();
    if ((int64_t)n > cap_) {  // the reference's array just grows (body_collection.go:273-291): so does the device image
        std::fprintf(stderr, "[INFO] device capacity %lld -> %lld bodies\n", (long long)cap_, (long long)(2 * n + 4096));
        recreate((int64_t)(2 * n + 4096), 0);
        stats_.regrows++;
    }
    grow(n);
    for (size_t i = 0; i < n; ++i) {
        const Body &b = *arr[i];
        x[i] = b.X; y[i] = b.Y; z[i] = b.Z; vx[i] = b.Vx; vy[i] = b.Vy; vz[i] = b.Vz;
        mass[i] = b.Mass; radius[i] = b.Radius; rest[i] = b.r; ff[i] = b.FragFactor; fs[i] = b.FragStep;
        beh[i] = (uint8_t)b.Behavior; flags[i] = flagsOf(b);
    }
    const int rc = nb_upload(h_, (int64_t)n, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data(),
                             mass.data(), radius.data(), rest.data(), ff.data(), fs.data(), beh.data(), flags.data());
    if (rc != NB_OK) std::fprintf(stderr, "[ERROR] nb_upload: %s\n", nb_last_error(h_));
    // nb_upload starts every body at fx = fy = fz = 0; a fragmenting body keeps applying the force of its
    // last Compute (body.go:152-155), which SyncToHost brought back into Body.fx,fy,fz
    bool anyFrag = false;
    for (size_t i = 0; i < n && !anyFrag; ++i) anyFrag = arr[i]->fragmenting;
    if (rc == NB_OK && anyFrag) {
        hfx.resize(n); hfy.resize(n); hfz.resize(n);
        for (size_t i = 0; i < n; ++i) { hfx[i] = arr[i]->fx; hfy[i] = arr[i]->fy; hfz[i] = arr[i]->fz; }
        if (nb_set_forces(h_, 0, (int64_t)n, hfx.data(), hfy.data(), hfz.data()) != NB_OK)
            std::fprintf(stderr, "[ERROR] nb_set_forces: %s\n", nb_last_error(h_));
    }
    n_ = (int64_t)n;
    dirty_ = false;
    hostStale_ = false;
    stats_.uploads++;
}

void GpuStepper::SyncToHost(BodyCollection &bc)
{
    if (!hostStale_) return;
    auto &arr = bc.GetArray();
    const size_t n = std::min(arr.size(), (size_t)n_);
    if (n == 0) { hostStale_ = false; return; }
    grow(n);
    const int rc = nb_download_state(h_, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data(), nullptr,
                                     nullptr, rest.data(), nullptr, flags.data());
    if (rc != NB_OK) { std::fprintf(stderr, "[ERROR] nb_download_state: %s\n", nb_last_error(h_)); return; }
    hfx.resize(n); hfy.resize(n); hfz.resize(n);
    const bool haveF = nb_get_forces(h_, hfx.data(), hfy.data(), hfz.data()) == NB_OK;
    for (size_t i = 0; i < n; ++i) {
        Body &b = *arr[i];
        b.X = x[i]; b.Y = y[i]; b.Z = z[i]; b.Vx = vx[i]; b.Vy = vy[i]; b.Vz = vz[i];
        b.r = rest[i];
        b.collided = false;
        if (haveF) { b.fx = hfx[i]; b.fy = hfy[i]; b.fz = hfz[i]; }
    }
    hostStale_ = false;
    stats_.downloads++;
}

bool GpuStepper::Step(BodyCollection &bc, double timeScaling, double R, ResultQueue &rq)
{
    auto &arr = bc.GetArray();
    // fragmenting bodies spawn their fragments on the host, as Body.Compute does (body.go:152-155)
    const bool inStep = !dirty_ && (int64_t)arr.size() == n_;
    // fragment() copies the body's CURRENT velocity into its fragments (fragcalc.go:97) and the device owns
    // the velocities between syncs: refresh the fragmenting bodies first (a few bodies: one small range
    // read each; many: one full sync)
    if (hostStale_ && inStep) {
        std::vector<size_t> fr;
        for (size_t i = 0; i < arr.size(); ++i)
            if (arr[i]->Exists && arr[i]->fragmenting) fr.push_back(i);
        if (fr.size() > 16) {
            SyncToHost(bc);
        } else {
            for (size_t i : fr) {
                double v[6];
                if (nb_download_state_range(h_, (int64_t)i, 1, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], nullptr, nullptr,
                                            nullptr, nullptr, nullptr) != NB_OK)
                    continue;
                Body &b = *arr[i];
                b.X = v[0]; b.Y = v[1]; b.Z = v[2]; b.Vx = v[3]; b.Vy = v[4]; b.Vz = v[5];
            }
        }
    }
    for (size_t i = 0; i < arr.size(); ++i) {
        Body &b = *arr[i];
        if (b.Exists && b.fragmenting) {
            b.fragment(bc);
            if (!b.Exists && inStep) {  // fully fragmented (fragcalc.go:114-116): one flag byte to the device
                const uint8_t fl1 = flagsOf(b);
                if (nb_patch(h_, (int64_t)i, 1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                             nullptr, nullptr, nullptr, nullptr, &fl1) != NB_OK)
                    dirty_ = true;
                stats_.patches++;
            }
        }
    }
    if (dirty_ || (int64_t)arr.size() != n_) upload(bc);
    const size_t n = arr.size();
    nb_step_result res{};
    int rc = nb_step(h_, timeScaling, R, NB_STEP_DEFAULT, &res);
    if (rc == NB_ERR_PAIR_OVERFLOW) {
        // the step was NOT applied: the device still holds the state of the cycle top.  The reference has
        // no event capacity (its list grows), so neither has the stepper: bring the state back, re-create
        // the handle with a larger event list and run the cycle again.
        std::fprintf(stderr, "[INFO] event capacity %lld exceeded: growing\n", (long long)pairCap_);
        hostStale_ = true;
        SyncToHost(bc);
        const int64_t bigger = std::max<int64_t>(4 * pairCap_, 16 * (int64_t)n + 65536);
        recreate(cap_, bigger);
        stats_.regrows++;
        upload(bc);
        rc = nb_step(h_, timeScaling, R, NB_STEP_DEFAULT, &res);
    }
    if (rc != NB_OK) {
        std::fprintf(stderr, "[ERROR] nb_step: %s\n", nb_last_error(h_));
        stats_.failed++;
        return false;
    }
    stats_.last = res;
    stats_.steps++;
    stats_.ms_device += res.ms_total;
    hostStale_ = true;
    grow(n);
    // The device resolved the whole event queue in the reference's order (ProcessMods): elastic
    // collisions, ResolveSubsume and the `fragmenting` flag.  The records keep the host objects in
    // step: a subsume is mirrored from the device's masses, a fragment decision runs the reference's
    // own handler for the host-only part (fragInfo).  Nothing is written back to the device.
    if (res.n_host_events > 0) {
        std::vector<nb_event> ev((size_t)res.n_host_events);
        int64_t m = 0;
        nb_get_host_events(h_, ev.data(), (int64_t)ev.size(), &m);
        int nFragInit = 0;
        bool anySubsume = false;
        for (int64_t k = 0; k < m; ++k) {
            nFragInit += ev[(size_t)k].kind == NB_EV_FRAG_INIT;
            anySubsume |= ev[(size_t)k].kind == NB_EV_SUBSUME && ev[(size_t)k].applied;
        }
        if (anySubsume &&
            nb_download_state(h_, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, mass.data(), nullptr, nullptr,
                              nullptr, flags.data()) != NB_OK) {
            std::fprintf(stderr, "[ERROR] nb_download_state: %s\n", nb_last_error(h_));
            return false;
        }
        // initiateFragmentation records where the body was while the queue was processed (fragcalc.go:77): the
        // position the cycle started from, not the one Update produced.  Many records: one bulk read.
        const bool bulkPos = nFragInit > 16;
        if (bulkPos && nb_get_cycle_top_positions(h_, 0, (int64_t)n, x.data(), y.data(), z.data()) != NB_OK) {
            std::fprintf(stderr, "[ERROR] nb_get_cycle_top_positions: %s\n", nb_last_error(h_));
            return false;
        }
        for (int64_t k = 0; k < m; ++k) {
            const nb_event &e = ev[(size_t)k];
            if (e.a < 0 || e.b < 0 || (size_t)e.a >= n || (size_t)e.b >= n) continue;
            if (e.kind == NB_EV_SUBSUME && e.applied) {
                Body &a = *arr[(size_t)e.a], &b = *arr[(size_t)e.b];
                if (b.Exists && !(flags[(size_t)e.b] & NB_F_EXISTS))  // body.go:243
                    std::fprintf(stderr, "[INFO] Body ID %d (mass %g) subsumed ID %d (mass %g)\n", a.Id, a.Mass, b.Id, b.Mass);
                a.Mass = mass[(size_t)e.a];
                b.Mass = mass[(size_t)e.b];
                if (!(flags[(size_t)e.a] & NB_F_EXISTS)) a.Exists = false;
                if (!(flags[(size_t)e.b] & NB_F_EXISTS)) b.Exists = false;
                stats_.subsumes++;
            } else if (e.kind == NB_EV_FRAG_INIT) {
                // in the reference's handling order (the library sorts them): a body named twice keeps the later call
                double p[3] = {x[(size_t)e.a], y[(size_t)e.a], z[(size_t)e.a]};
                if (!bulkPos && nb_get_cycle_top_positions(h_, e.a, 1, &p[0], &p[1], &p[2]) != NB_OK) continue;
                arr[(size_t)e.a]->initiateFragmentationAt(e.f1, e.dist, p[0], p[1], p[2]);
            }
        }
    }
    // Renderables from the float32 snapshot (13 B/body) — computation-runner.go:317-320
    const float *rxyz = pinXyz_;
    const uint8_t *rex = pinExists_;
    if (!rxyz) {  // fallback: explicit copy
        if (n > 0 && nb_download_render(h_, xyz.data(), exists.data()) != NB_OK) {
            std::fprintf(stderr, "[ERROR] nb_download_render: %s\n", nb_last_error(h_));
            return false;
        }
        rxyz = xyz.data();
        rex = exists.data();
    }
    rq.queue.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        Body &b = *arr[i];
        Renderable r;
        r.Id = b.Id;
        if (b.Exists && !rex[i]) {
            std::fprintf(stderr, "[ERROR] NaN values. id=%d (removing from sim)\n", b.Id);  // body.go:135
            b.Exists = false;
        }
        if (b.Exists) {
            r.Exists = true;
            r.X = rxyz[3 * i]; r.Y = rxyz[3 * i + 1]; r.Z = rxyz[3 * i + 2];
            r.Radius = b.Radius; r.IsSun = b.IsSun; r.Intensity = (float)b.intensity; r.Color = b.Color;
        }
        rq.Add(r);
    }
    return true;
}

void GpuStepper::Reserve(BodyCollection &bc, int64_t count)
{
    if (count <= cap_) return;
    SyncToHost(bc);  // while host and device indices still correspond
    dirty_ = true;   // upload() re-creates the handle with room for the new count
}

// After BodyCollection.Cycle: keep the device array in step with the host array without a full
// re-upload when only deaths (stable compaction) and/or adds (append) happened.
void GpuStepper::AfterCycle(BodyCollection &bc, bool arrayChanged, int64_t newCount, double R)
{
    if (!arrayChanged || dirty_) return;
    auto &arr = bc.GetArray();
    int64_t n_dev = n_;
    // deaths (NaN cull on the device, subsume / delete patched from the host) are compacted on the
    // device exactly like Cycle compacted the host array; a no-op pass when nothing died
    if (nb_compact(h_, &n_dev, nullptr, 0) != NB_OK) { dirty_ = true; return; }
    if (n_dev != n_) stats_.compacts++;
    const int64_t adds = newCount - n_dev;
    if (adds < 0 || newCount > cap_) { dirty_ = true; return; }
    if (adds > 0) {
        grow((size_t)adds);
        for (int64_t k = 0; k < adds; ++k) {
            const Body &b = *arr[(size_t)(n_dev + k)];
            x[k] = b.X; y[k] = b.Y; z[k] = b.Z; vx[k] = b.Vx; vy[k] = b.Vy; vz[k] = b.Vz;
            mass[k] = b.Mass; radius[k] = b.Radius; ff[k] = b.FragFactor; fs[k] = b.FragStep;
            beh[k] = (uint8_t)b.Behavior; flags[k] = flagsOf(b);
        }
        if (nb_append(h_, adds, R, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data(), mass.data(),
                      radius.data(), ff.data(), fs.data(), beh.data(), flags.data()) != NB_OK) {
            dirty_ = true;
            return;
        }
        stats_.appends++;
    }
    n_ = newCount;
}

// ---------------------------------------------------------------- ComputationRunner
ComputationRunner::ComputationRunner(int workerCnt, double timeScaling, bool /*barnesHut*/, ResultQueueHolder *rqh,
                                     BodyCollection *bc, int device, int64_t capacity)
    : workerCnt_(workerCnt), bc_(bc), timeScaling_(timeScaling), rqh_(rqh)
{
    if (capacity <= 0) capacity = std::max<int64_t>(4096, 2 * (int64_t)bc->Count() + 4096);
    stepper_ = std::make_unique<GpuStepper>(device, capacity);
    bc_->syncFromDevice = [this] { stepper_->SyncToHost(*bc_); };
}

ComputationRunner::~ComputationRunner()
{
    Stop();
    bc_->syncFromDevice = nullptr;
}

ComputationRunner &ComputationRunner::SetMaxIterations(int maxIteration)
{
    maxIteration_ = maxIteration;
    return *this;
}

ComputationRunner &ComputationRunner::Start()
{
    stop_ = false;
    running_ = true;
    th_ = std::thread([this] { run(); });
    return *this;
}

void ComputationRunner::Stop()
{
    stop_ = true;
    if (th_.joinable()) th_.join();
}

void ComputationRunner::SetWorkers(int workerCnt) { workerCnt_ = workerCnt; }  // no pool to resize

void ComputationRunner::SetTimeScaling(double ts)
{
    std::lock_guard<std::mutex> g(ctl_);
    pendingTs_ = ts;
    haveTs_ = true;
}

void ComputationRunner::SetCoefficientOfRestitution(double R)
{
    std::lock_guard<std::mutex> g(ctl_);
    pendingR_ = R;
    haveR_ = true;
}

void ComputationRunner::RemoveBodies(int deletes)
{
    std::lock_guard<std::mutex> g(ctl_);
    pendingDel_ = deletes;
    haveDel_ = true;
}

// computation-runner.go:176-216
void ComputationRunner::processDeletes()
{
    int delCnt;
    {
        std::lock_guard<std::mutex> g(ctl_);
        if (!haveDel_) return;
        delCnt = pendingDel_;
        haveDel_ = false;
    }
    stepper_->SyncToHost(*bc_);
    int removedCnt = 0;
    if (delCnt == -1) {
        bc_->IterateOnce([&](Body &b) {
            if (b.Exists) { b.SetNotExists(); removedCnt++; }
        });
    } else if (delCnt > 0) {
        const int count = bc_->Count();
        const int step = delCnt > count ? 1 : count / delCnt;
        int iter = 0;
        bool shouldRemove = false;
        bc_->IterateOnce([&](Body &b) {
            if (iter % step == 0) shouldRemove = true;
            iter++;
            // the reference keeps iterating after removedCnt reaches delCnt (`return` only leaves the
            // closure), so more than delCnt bodies can be removed; restated as is (:198-208)
            if (shouldRemove && !b.Pinned && b.Exists) {
                b.SetNotExists();
                shouldRemove = false;
                removedCnt++;
            }
        });
    }
    if (removedCnt) stepper_->MarkDirty();
    std::fprintf(stderr, "[INFO] Computation runner set %d bodies to not exist\n", removedCnt);
}

void ComputationRunner::runOneComputation()
{
    iterations_++;
    {
        std::lock_guard<std::mutex> g(ctl_);
        if (haveTs_) { timeScaling_ = pendingTs_; haveTs_ = false; }
        if (haveR_) { R_ = pendingR_; haveR_ = false; }
    }
    processDeletes();
    bc_->HandleGetBody();
    if (bc_->HandleModBody()) stepper_->MarkDirty();
    auto [rq, ok] = rqh_->NewResultQueue();
    if (!ok) {
        skipped_++;
        std::this_thread::sleep_for(std::chrono::milliseconds(5));  // :276-279
        return;
    }
    if (bc_->Count() == 0 && bc_->pendingAdds() == 0) {
        std::this_thread::sleep_for(std::chrono::milliseconds(5));  // no bodies (:313)
    }
    if (!stepper_->Step(*bc_, timeScaling_, R_, *rq)) {
        // a failed device step changed nothing: publish nothing and do not Cycle (the queue handed out by
        // NewResultQueue is simply dropped); the error is on stderr and in stats().failed
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
        return;
    }
    rqh_->Add(rq);
    stepper_->Reserve(*bc_, (int64_t)bc_->Count() + (int64_t)bc_->pendingAdds());
    const bool changed = bc_->Cycle(R_);
    stepper_->AfterCycle(*bc_, changed, bc_->Count(), R_);
    computations_++;
}

void ComputationRunner::run()
{
    startTime_ = std::chrono::steady_clock::now();
    while (!stop_) {
        runOneComputation();
        if (maxIteration_ > 0 && --maxIteration_ == 0) break;
        std::this_thread::yield();
    }
    stepper_->SyncToHost(*bc_);
    stopTime_ = std::chrono::steady_clock::now();
    running_ = false;
}

void ComputationRunner::PrintStats()
{
    const double totalMillis = std::chrono::duration<double, std::milli>(stopTime_ - startTime_).count();
    const double fps = totalMillis > 0 ? (double)computations_ / totalMillis * 1000 : 0;
    const StepStats &s = stepper_->stats();
    std::printf("Runner\n workerCnt: %d (GPU path: no worker pool)\n iterations: %llu\n computations: %llu\n"
                " skipped (no queue capacity): %llu\n device ms/computation: %g\n frames per second: %g\n"
                " elapsed time: %gs\nGpuStepper\n uploads: %llu appends: %llu compacts: %llu downloads: %llu patches: %llu subsumes: %llu\n",
                workerCnt_, (unsigned long long)iterations_, (unsigned long long)computations_,
                (unsigned long long)skipped_, s.steps ? s.ms_device / (double)s.steps : 0.0, fps, totalMillis / 1000,
                (unsigned long long)s.uploads, (unsigned long long)s.appends, (unsigned long long)s.compacts,
                (unsigned long long)s.downloads, (unsigned long long)s.patches, (unsigned long long)s.subsumes);
}

}  // namespace nbodygo
