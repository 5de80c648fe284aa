// nbody_host.h — host side above the C ABI, in C++ because the reference's host is compiled
// code (Go) and this image has no Go toolchain.  It mirrors the reference's interface for the
// path around the compute cycle — same names, argument meaning and error behaviour — so that a
// reader of cmd/body, cmd/runner and cmd/sim finds every entry point again:
//
//   globals      cmd/globals/globals.go          CollisionBehavior, BodyColor, Parse*, SafeParseFloat
//   Body         cmd/body/body.go                fields, NewBody, SetNotExists, SetSun, ApplyMods, NextId,
//                                                ResolveSubsume; fragcalc.go: doFragment / fragment
//   Renderable   cmd/body/renderable.go
//   BodyCollection cmd/body/body_collection.go   Enqueue, ProcessMods, Cycle, GetBody/HandleGetBody,
//                                                ModBody/HandleModBody, IterateOnce, Count
//   ResultQueueHolder cmd/runner/resultqueue.go  NewResultQueue, Add, Next, Resize, MaxQueues
//   ComputationRunner cmd/runner/computation-runner.go  Start/Stop, SetWorkers, SetTimeScaling,
//                                                SetCoefficientOfRestitution, RemoveBodies, PrintStats
//   GpuStepper   (new) replaces WorkPool + the block computation-runner.go:285-320 through
//                include/nbody_b200.h — the C++ twin of the cgo shim in INTEGRATION.md
//   sim          cmd/sim/fromcsv.go, simgen.go   FromCsv, seeded Sim1..Sim5 / SimTest, headless run
//
// What is NOT here: rendering, gRPC, Prometheus, logging filters (out of scope).
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/nbody_b200.h"

namespace nbodygo {

// ---------------------------------------------------------------- globals (cmd/globals/globals.go)
enum CollisionBehavior : int { None = 0, Subsume = 1, Elastic = 2, Fragment = 3 };
enum BodyColor : int { Random = 0, Black, White, Darkgray, Gray, Lightgray, Red, Green, Blue, Yellow, Magenta, Cyan,
                       Orange, Brown, Pink };
CollisionBehavior ParseCollisionBehavior(const std::string &s);  // unknown → Elastic   (globals.go:39-46)
bool ParseBoolean(const std::string &s);                          // t,true,1,y,yes      (globals.go:51-57)
BodyColor ParseBodyColor(const std::string &s);                   // unknown → Random    (globals.go:62-70)
double SafeParseFloat(const std::string &s, double cur);          // parse error → cur   (globals.go:97-102)

// ---------------------------------------------------------------- Body (cmd/body/body.go:34-52)
class BodyCollection;
struct Body;
using BodyPtr = std::shared_ptr<Body>;

struct FragInfo {  // body.go:26-31
    double radius = 0, newRadius = 0, mass = 0;
    int fragments = 0;
    double x = 0, y = 0, z = 0;
};

struct Body {
    int Id = 0;
    std::string Name, Class;
    double X = 0, Y = 0, Z = 0, Vx = 0, Vy = 0, Vz = 0, Radius = 0, Mass = 0;
    double FragFactor = 0, FragStep = 0;
    CollisionBehavior Behavior = Elastic;  // Body.CollisionBehavior
    BodyColor Color = Random;              // Body.BodyColor
    bool IsSun = false, Exists = true, WithTelemetry = false, Pinned = false;
    // unexported in Go
    double r = 1;  // coefficient of restitution in force for this body
    bool fragmenting = false;
    double intensity = 0;
    FragInfo fragInfo;
    double fx = 0, fy = 0, fz = 0;
    bool collided = false;

    void SetNotExists();                            // body.go:93-96
    void SetSun(double intensity_);                 // body.go:101-104
    bool ApplyMods(const std::vector<std::string> &mods);  // body.go:274-313
    void ResolveSubsume(Body &other);               // body.go:228-244
    // fragcalc.go:54-83: called with the factors the device computed (NB_EV_FRAGMENT)
    void DoFragment(Body &other, double thisFactor, double otherFactor);
    void initiateFragmentation(double fragFactor);
    // the same with the Mass and position the body had when the event was handled (NB_EV_FRAG_INIT record +
    // nb_get_cycle_top_positions): the device resolves the queue, the host only keeps fragInfo
    void initiateFragmentationAt(double fragFactor, double massThen, double x, double y, double z);
    // fragcalc.go:90-117: spawns up to maxFragsPerCycle+1 bodies per cycle into bc (as add events)
    void fragment(BodyCollection &bc);
};

// body.go:56-89
BodyPtr NewBody(int id, double x, double y, double z, double vx, double vy, double vz, double mass, double radius,
                CollisionBehavior behavior, BodyColor color, double fragFactor, double fragStep, bool withTelemetry,
                const std::string &name, const std::string &cls, bool pinned);
int NextId();          // body.go:316-329
void ResetIdGenerator();  // tests only
void SetNextId(int id);    // resume from a state dump: NextId() continues where the dumped run stopped

// ---------------------------------------------------------------- Renderable (cmd/body/renderable.go)
struct Renderable {
    int Id = 0;
    bool Exists = false;
    float X = 0, Y = 0, Z = 0;
    double Radius = 0;
    bool IsSun = false;
    float Intensity = 0;
    BodyColor Color = Random;
};

// ---------------------------------------------------------------- events (cmd/body/event.go)
enum class EventType { Collision, Subsume, Add, Fragment };
struct Event {
    EventType type;
    BodyPtr b1, b2;          // twoBodies
    BodyPtr b;               // eneBody (add)
    double f1 = 0, f2 = 0;   // fragment factors
};
Event NewAdd(BodyPtr b);                                      // event.go:103-111
Event newSubsume(BodyPtr b1, BodyPtr b2);                     // event.go:90-98
Event newFragment(BodyPtr b1, BodyPtr b2, double f1, double f2);

enum class ModBodyResult { NoMatch, ModNone, ModSome, ModAll };  // cmd/grpcsimcb

// ---------------------------------------------------------------- BodyCollection (body_collection.go)
class BodyCollection {
public:
    explicit BodyCollection(const std::vector<BodyPtr> &bodies);  // NewSimBodyCollection :45-68
    std::vector<BodyPtr> &GetArray() { return arr_; }              // :70-72
    // Deferred events. Unlike the reference's 1000-slot channel (:82-88) nothing is dropped.
    void Enqueue(const Event &ev);
    // Handles subsume / fragment events (collision events are resolved on the device) :212-233
    void ProcessMods();
    // Removes !Exists bodies (stable) and appends enqueued adds with r = R :253-296.
    // Returns true if the array changed.
    bool Cycle(double R);
    int Count();
    void IterateOnce(const std::function<void(Body &)> &c);  // :196-200
    // rendezvous with the runner thread (:111-191): block until Handle* services the request
    BodyPtr GetBody(int id, const std::string &name);
    void HandleGetBody();
    ModBodyResult ModBody(int id, const std::string &name, const std::string &cls,
                          const std::vector<std::string> &mods);
    bool HandleModBody();  // returns true if a request was serviced (device state is then stale)
    int cycle() const { return cycle_; }
    void setCycle(int c) { cycle_ = c; }  // resume from a state dump (the cycle number seeds fragment())
    int pendingAdds();
    // set by the runner: brings host bodies up to date with the device before they are read
    std::function<void()> syncFromDevice;

private:
    std::vector<BodyPtr> arr_;
    std::list<Event> events_;
    std::mutex lock_;
    int cycle_ = 0;
    struct GetReq { int id; std::string name; };
    struct ModReq { int id; std::string name, cls; std::vector<std::string> mods; };
    std::mutex chLock_;
    std::condition_variable chCv_;
    std::deque<GetReq> getBodyCh_;
    std::deque<BodyPtr> sendBodyCh_;
    bool sendReady_ = false;
    std::deque<ModReq> modBodyCh_;
    std::deque<ModBodyResult> modBodyResultCh_;
};

// ---------------------------------------------------------------- result queues (resultqueue.go)
struct ResultQueue {
    unsigned QueueNum = 0;
    std::vector<Renderable> queue;
    void Add(const Renderable &r) { queue.push_back(r); }
    const std::vector<Renderable> &Queue() const { return queue; }
};
using ResultQueuePtr = std::shared_ptr<ResultQueue>;

class ResultQueueHolder {
public:
    explicit ResultQueueHolder(int maxQueues);      // NewResultQueueHolder :75-93
    std::pair<ResultQueuePtr, bool> NewResultQueue();  // :101-118  (nullptr,false) when full → cycle skipped
    void Add(ResultQueuePtr q);                      // :58-70 (fatal when over physical capacity)
    std::pair<ResultQueuePtr, bool> Next();          // :124-136
    int MaxQueues();                                 // :141-146
    bool Resize(int maxQueues);                      // :168-198
    int Len();

private:
    std::mutex lock_;
    std::deque<ResultQueuePtr> ch_;
    int maxQueues_, physCap_;
    unsigned queueNum_ = 0;
};

// ---------------------------------------------------------------- GPU stepper (replaces WorkPool)
struct StepStats {
    nb_step_result last{};
    uint64_t steps = 0, uploads = 0, appends = 0, compacts = 0, downloads = 0, patches = 0, subsumes = 0, regrows = 0,
             failed = 0;
    double ms_device = 0;
};

class GpuStepper {
public:
    GpuStepper(int device, int64_t capacity);  // NewWorkPool's place (workpool.go:129-142); throws on failure
    ~GpuStepper();
    // One compute cycle on the device + Renderables into rq (computation-runner.go:285-320).
    // Returns false (with [ERROR] logged) if the device step failed.
    bool Step(BodyCollection &bc, double timeScaling, double R, ResultQueue &rq);
    void MarkDirty() { dirty_ = true; }        // host bodies changed: re-upload before the next step
    void SyncToHost(BodyCollection &bc);       // device state → host bodies (GetBody, mods, end of run)
    void AfterCycle(BodyCollection &bc, bool arrayChanged, int64_t newCount, double R);
    // The reference's array simply grows (body_collection.go:273-291).  Call before Cycle appends: when
    // `count` bodies will not fit, the host bodies are refreshed from the device while indices still match,
    // and the next step re-creates the handle with room to spare and re-uploads.
    void Reserve(BodyCollection &bc, int64_t count);
    int64_t Capacity() const { return cap_; }
    const StepStats &stats() const { return stats_; }
    nb_handle handle() { return h_; }

private:
    void upload(BodyCollection &bc);
    void grow(size_t n);
    void recreate(int64_t capacity, int64_t pairCapacity);  // throws on failure, like the constructor
    nb_handle h_ = nullptr;
    int device_ = 0;
    int64_t n_ = 0, cap_ = 0, pairCap_ = 0;
    std::vector<double> hfx, hfy, hfz;
    bool dirty_ = true, hostStale_ = false;
    std::vector<double> x, y, z, vx, vy, vz, mass, radius, rest, ff, fs;
    std::vector<uint8_t> beh, flags, exists;
    std::vector<float> xyz;
    float *pinXyz_ = nullptr;        // library-owned pinned snapshot buffers (nb_render_buffers)
    uint8_t *pinExists_ = nullptr;
    StepStats stats_;
};

// ---------------------------------------------------------------- ComputationRunner
class ComputationRunner {
public:
    // NewComputationRunner (computation-runner.go:81-98). workerCnt and barnesHut are accepted for
    // interface compatibility: the GPU path has no worker pool and is always brute force.
    ComputationRunner(int workerCnt, double timeScaling, bool barnesHut, ResultQueueHolder *rqh, BodyCollection *bc,
                      int device = 0, int64_t capacity = 0);
    ~ComputationRunner();
    ComputationRunner &SetMaxIterations(int maxIteration);  // :102-105
    ComputationRunner &Start();                             // :108-111
    void Stop();                                            // :114-120
    void SetWorkers(int workerCnt);                         // :124-128 (no-op on the GPU path)
    int WorkerCount() const { return workerCnt_; }
    double TimeScaling() const { return timeScaling_; }
    void SetTimeScaling(double ts);                         // :136-138
    double CoefficientOfRestitution() const { return R_; }
    void SetCoefficientOfRestitution(double R);             // :156-158
    void RemoveBodies(int deletes);                         // :171-173
    void PrintStats();                                      // :61-71
    void runOneComputation();                               // :267-326 (public for single-threaded tests)
    bool Running() const { return running_; }
    uint64_t Computations() const { return computations_; }
    uint64_t Iterations() const { return iterations_; }
    GpuStepper &Stepper() { return *stepper_; }

private:
    void run();
    void processDeletes();
    int workerCnt_;
    std::atomic<bool> stop_{false}, running_{false};
    uint64_t iterations_ = 0, computations_ = 0, skipped_ = 0;
    std::chrono::steady_clock::time_point startTime_, stopTime_;
    std::unique_ptr<GpuStepper> stepper_;
    BodyCollection *bc_;
    int maxIteration_ = 0;
    double timeScaling_;
    ResultQueueHolder *rqh_;
    double R_ = 1;
    std::mutex ctl_;
    bool haveTs_ = false, haveR_ = false, haveDel_ = false;
    double pendingTs_ = 0, pendingR_ = 0;
    int pendingDel_ = 0;
    std::thread th_;
};

// ---------------------------------------------------------------- sim (cmd/sim)
// FromCsv (fromcsv.go:46-116): 8 required float fields, optional is_sun, collision, color,
// frag_factor, frag_step; '#' comments; rows with parse errors are skipped.
std::vector<BodyPtr> FromCsv(const std::string &csvPath, int bodyCount, CollisionBehavior defaultBehavior,
                             BodyColor defaultColor);
bool WriteCsv(const std::string &csvPath, const std::vector<BodyPtr> &bodies);  // %.17g, same 13 columns
// Seeded restatements of the generators of simgen.go (the reference seeds from the clock).
std::vector<BodyPtr> Generate(const std::string &simName, int bodyCount, CollisionBehavior behavior, BodyColor color,
                              const std::string &simArgs, uint64_t seed);

// What the reference's CSV cannot carry (it is an INPUT format: fromcsv.go:15-47): identity (Id, Name, Class, Pinned,
// WithTelemetry) and the unexported running state of a Body (r, fragmenting, fragInfo, fx fy fz, intensity), plus the
// id generator, the collection's cycle counter and the runner's R.  Written next to the final-state CSV
// (`nbody_server --dump-final-state`), read back with `--resume-state`: together they are a bit-exact checkpoint
// for any R and with fragmentation in flight (tests/test_checkpoint.py).  One line per body, array order, floats as
// C99 hex (%a).
struct RunState {
    int nextId = -1, cycle = -1;  // -1: leave the id generator / cycle counter as they are
    double R = 1;
};
bool WriteState(const std::string &path, const std::vector<BodyPtr> &bodies, const RunState &rs);
// applies the sidecar to bodies freshly read by FromCsv (same count, same order); false on any mismatch
bool ReadState(const std::string &path, std::vector<BodyPtr> &bodies, RunState &rs);

struct HeadlessResult {
    uint64_t computations = 0, iterations = 0;
    double seconds = 0, fps = 0, interactionsPerSec = 0;
    int finalBodies = 0;
    std::vector<BodyPtr> bodies;  // the collection's array when the run stopped (fragments and added bodies included)
    RunState state;  // id generator, cycle counter and R when the run stopped
};
// nBodySim.Run with render == false (nbodysim.go:78-133): start the runner, drain the result queues,
// stop after runMillis (or maxIterations if > 0), print stats.
HeadlessResult RunHeadless(std::vector<BodyPtr> bodies, double timeScaling, int runMillis, int maxIterations,
                           int device, bool quiet, const RunState *start = nullptr);

}  // namespace nbodygo
