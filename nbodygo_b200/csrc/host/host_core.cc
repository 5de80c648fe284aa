// host_core.cc — globals, Body, events, BodyCollection, ResultQueueHolder (see nbody_host.h).
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "nbody_host.h"

namespace nbodygo {

// ---------------------------------------------------------------- globals
static std::string lower(std::string s)
{
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}

CollisionBehavior ParseCollisionBehavior(const std::string &s)
{
    static const char *names[] = {"none", "subsume", "elastic", "fragment"};
    const std::string l = lower(s);
    for (int i = 0; i < 4; ++i)
        if (l == names[i]) return (CollisionBehavior)i;
    return Elastic;
}

bool ParseBoolean(const std::string &s)
{
    const std::string l = lower(s);
    return l == "t" || l == "true" || l == "1" || l == "y" || l == "yes";
}

BodyColor ParseBodyColor(const std::string &s)
{
    static const char *names[] = {"random", "black", "white", "darkgray", "gray", "lightgray", "red", "green",
                                  "blue", "yellow", "magenta", "cyan", "orange", "brown", "pink"};
    const std::string l = lower(s);
    for (int i = 0; i < 15; ++i)
        if (l == names[i]) return (BodyColor)i;
    return Random;
}

double SafeParseFloat(const std::string &s, double cur)
{
    if (s.empty()) return cur;
    char *end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end != '\0') return cur;
    return v;
}

// ---------------------------------------------------------------- Body
static const double kFourThirdsPi = 3.14159265358979323846 * (4 / 3);  // body.go:20 — integer division, as in Go
static const double kFourPi = 3.14159265358979323846 * 4;
static const int kMaxFragsPerCycle = 100;
static const double kMaxFrags = 2000;

BodyPtr NewBody(int id, double x, double y, double z, double vx, double vy, double vz, double mass, double radius,
                CollisionBehavior behavior, BodyColor color, double fragFactor, double fragStep, bool withTelemetry,
                const std::string &name, const std::string &cls, bool pinned)
{
    auto b = std::make_shared<Body>();
    b->Id = id; b->Name = name; b->Class = cls;
    b->X = x; b->Y = y; b->Z = z; b->Vx = vx; b->Vy = vy; b->Vz = vz;
    b->Radius = radius; b->Mass = mass;
    b->FragFactor = fragFactor; b->FragStep = fragStep;
    b->Behavior = behavior; b->Color = color;
    b->r = 1; b->Exists = true; b->WithTelemetry = withTelemetry; b->Pinned = pinned;
    return b;
}

void Body::SetNotExists()
{
    Mass = 0;
    Exists = false;
}

void Body::SetSun(double intensity_)
{
    IsSun = true;
    intensity = intensity_;
}

bool Body::ApplyMods(const std::vector<std::string> &mods)
{
    for (const auto &mod : mods) {
        // strings.Split(mod, "=") must yield exactly two parts
        const size_t eq = mod.find('=');
        if (eq == std::string::npos || mod.find('=', eq + 1) != std::string::npos) continue;
        std::string key = mod.substr(0, eq);
        const std::string val = mod.substr(eq + 1);
        std::transform(key.begin(), key.end(), key.begin(), [](unsigned char c) { return (char)std::toupper(c); });
        if (key == "X") X = SafeParseFloat(val, X);
        else if (key == "Y") Y = SafeParseFloat(val, Y);
        else if (key == "Z") Z = SafeParseFloat(val, Z);
        else if (key == "VX") Vx = SafeParseFloat(val, Vx);
        else if (key == "VY") Vy = SafeParseFloat(val, Vy);
        else if (key == "VZ") Vz = SafeParseFloat(val, Vz);
        else if (key == "MASS") Mass = SafeParseFloat(val, Mass);
        else if (key == "RADIUS") Radius = SafeParseFloat(val, Radius);
        else if (key == "FRAG-FACTOR") FragFactor = SafeParseFloat(val, FragFactor);
        else if (key == "FRAG-STEP") FragStep = SafeParseFloat(val, FragStep);
        else if (key == "COLLISION") Behavior = ParseCollisionBehavior(val);
        else if (key == "COLOR") { if (!IsSun) Color = ParseBodyColor(val); }
        else if (key == "TELEMETRY") WithTelemetry = ParseBoolean(val);
        else if (key == "EXISTS") Exists = ParseBoolean(val);
    }
    return true;
}

void Body::ResolveSubsume(Body &other)
{
    const double thisMass = Mass, otherMass = other.Mass;
    Mass = thisMass + otherMass;
    other.SetNotExists();
}

void Body::DoFragment(Body &other, double thisFactor, double otherFactor)
{
    if (Behavior == Fragment && thisFactor > FragFactor) initiateFragmentation(thisFactor);
    if (other.Behavior == Fragment && otherFactor > other.FragFactor) other.initiateFragmentation(otherFactor);
}

void Body::initiateFragmentation(double fragFactor)
{
    const double fragDelta = FragFactor > 10 ? 10 : fragFactor - FragFactor;
    const double fragments = std::fmin(fragDelta * FragStep, kMaxFrags);
    if (fragments <= 1) {
        Behavior = Fragment;
        return;
    }
    fragmenting = true;
    const double volume = kFourThirdsPi * Radius * Radius * Radius;
    // math.Pow(x, 1/3): 1/3 is integer division in the reference (== 0) so Pow(...) == 1 and the
    // new radius is max(1, .1) == 1; restated faithfully (fragcalc.go:80)
    const double newRadius = std::fmax(std::pow(((volume / fragments) * 3) / kFourPi, (double)(1 / 3)), .1);
    const double newMass = Mass / fragments;
    fragInfo = FragInfo{Radius, newRadius, newMass, (int)fragments, X, Y, Z};
}

void Body::initiateFragmentationAt(double fragFactor, double massThen, double x, double y, double z)
{
    // run the reference's own routine on the state of that moment, then put the current state back
    const double m = Mass, px = X, py = Y, pz = Z;
    Mass = massThen; X = x; Y = y; Z = z;
    initiateFragmentation(fragFactor);
    Mass = m; X = px; Y = py; Z = pz;
}

// util.GetVectorEven (cmd/util/vectorutil.go:32-42), seeded per body instead of from the clock
static void vectorEven(std::mt19937_64 &rng, double cx, double cy, double cz, double radius, double out[3])
{
    std::uniform_real_distribution<double> u(0.0, 1.0);
    double x, y, z, d = 2;
    while (d > 1) {
        x = u(rng) * 2 - 1;
        y = u(rng) * 2 - 1;
        z = u(rng) * 2 - 1;
        d = x * x + y * y + z * z;
    }
    out[0] = x * radius + cx;
    out[1] = y * radius + cy;
    out[2] = z * radius + cz;
}

void Body::fragment(BodyCollection &bc)
{
    // deterministic stand-in for the reference's clock seed; the cycle number keeps a body that is
    // re-initiated every cycle from dropping its fragments on top of the previous cycle's
    std::mt19937_64 rng(0x9E3779B97F4A7C15ull ^ ((uint64_t)Id << 40) ^ ((uint64_t)bc.cycle() << 16) ^
                        (uint64_t)fragInfo.fragments);
    int cnt = 0;
    while (fragInfo.fragments > 0) {
        fragInfo.fragments--;
        double v[3];
        vectorEven(rng, fragInfo.x, fragInfo.y, fragInfo.z, fragInfo.radius * .9, v);
        auto toAdd = std::make_shared<Body>();
        toAdd->Id = NextId();
        toAdd->Name = Name; toAdd->Class = Class;
        toAdd->X = v[0]; toAdd->Y = v[1]; toAdd->Z = v[2];
        toAdd->Vx = Vx; toAdd->Vy = Vy; toAdd->Vz = Vz;
        toAdd->Mass = fragInfo.mass; toAdd->Radius = fragInfo.newRadius;
        toAdd->Behavior = Elastic; toAdd->Color = Color;
        toAdd->Exists = true;
        toAdd->r = 0;  // Go zero value: the literal in fragcalc.go:95-103 does not set r; Cycle sets it to R
        bc.Enqueue(NewAdd(toAdd));
        if (++cnt > kMaxFragsPerCycle) break;
    }
    if (fragInfo.fragments <= 0) Exists = false;
}

static std::mutex g_idLock;
static int g_id = 0;
int NextId()
{
    std::lock_guard<std::mutex> g(g_idLock);
    return g_id++;
}
void ResetIdGenerator()
{
    std::lock_guard<std::mutex> g(g_idLock);
    g_id = 0;
}
void SetNextId(int id)
{
    std::lock_guard<std::mutex> g(g_idLock);
    g_id = id;
}

// ---------------------------------------------------------------- events
Event NewAdd(BodyPtr b) { return Event{EventType::Add, nullptr, nullptr, std::move(b)}; }
Event newSubsume(BodyPtr b1, BodyPtr b2) { return Event{EventType::Subsume, std::move(b1), std::move(b2), nullptr}; }
Event newFragment(BodyPtr b1, BodyPtr b2, double f1, double f2)
{
    return Event{EventType::Fragment, std::move(b1), std::move(b2), nullptr, f1, f2};
}

// ---------------------------------------------------------------- BodyCollection
BodyCollection::BodyCollection(const std::vector<BodyPtr> &bodies) : arr_(bodies) {}

void BodyCollection::Enqueue(const Event &ev)
{
    std::lock_guard<std::mutex> g(lock_);
    events_.push_front(ev);  // handleEvents: PushFront (body_collection.go:99-101)
}

void BodyCollection::ProcessMods()
{
    std::vector<Event> evs;
    {
        std::lock_guard<std::mutex> g(lock_);
        for (auto it = events_.begin(); it != events_.end();) {
            if (it->type != EventType::Add) {
                evs.push_back(*it);
                it = events_.erase(it);
            } else {
                ++it;
            }
        }
    }
    for (auto &e : evs) {  // Front→Next order == reverse arrival
        if (e.type == EventType::Subsume) e.b1->ResolveSubsume(*e.b2);
        else if (e.type == EventType::Fragment) e.b1->DoFragment(*e.b2, e.f1, e.f2);
    }
}

int BodyCollection::pendingAdds()
{
    std::lock_guard<std::mutex> g(lock_);
    int c = 0;
    for (auto &e : events_) c += e.type == EventType::Add;
    return c;
}

bool BodyCollection::Cycle(double R)
{
    size_t cnt = 0;
    for (auto &b : arr_) cnt += b->Exists;
    std::lock_guard<std::mutex> g(lock_);
    bool changed = false;
    if (cnt < arr_.size()) {
        std::vector<BodyPtr> out;
        out.reserve(cnt + events_.size());
        for (auto &b : arr_)
            if (b->Exists) out.push_back(b);
        arr_.swap(out);
        changed = true;
    }
    // adds in list order (Front→Next), each with r = R
    for (auto &e : events_) {
        if (e.type == EventType::Add) {
            e.b->r = R;
            arr_.push_back(e.b);
            changed = true;
        }
    }
    events_.clear();  // bc.events.Init(): pending non-add events are wiped, as in the reference
    cycle_++;
    return changed;
}

int BodyCollection::Count()
{
    std::lock_guard<std::mutex> g(lock_);
    return (int)arr_.size();
}

void BodyCollection::IterateOnce(const std::function<void(Body &)> &c)
{
    for (size_t i = 0, n = arr_.size(); i < n; ++i) c(*arr_[i]);
}

BodyPtr BodyCollection::GetBody(int id, const std::string &name)
{
    std::unique_lock<std::mutex> g(chLock_);
    getBodyCh_.push_back({id, name});
    chCv_.wait(g, [&] { return !sendBodyCh_.empty(); });
    BodyPtr b = sendBodyCh_.front();
    sendBodyCh_.pop_front();
    return b;
}

void BodyCollection::HandleGetBody()
{
    GetReq req;
    {
        std::lock_guard<std::mutex> g(chLock_);
        if (getBodyCh_.empty()) return;
        req = getBodyCh_.front();
        getBodyCh_.pop_front();
    }
    if (syncFromDevice) syncFromDevice();
    BodyPtr found;
    for (auto &b : arr_) {
        if ((!req.name.empty() && req.name == b->Name) || req.id == b->Id) {
            // clone (doSendBody :138-149): the caller never sees the live object
            found = NewBody(b->Id, b->X, b->Y, b->Z, b->Vx, b->Vy, b->Vz, b->Mass, b->Radius, b->Behavior, b->Color,
                            b->FragFactor, b->FragStep, b->WithTelemetry, b->Name, b->Class, b->Pinned);
            break;
        }
    }
    std::lock_guard<std::mutex> g(chLock_);
    sendBodyCh_.push_back(found);
    chCv_.notify_all();
}

ModBodyResult BodyCollection::ModBody(int id, const std::string &name, const std::string &cls,
                                      const std::vector<std::string> &mods)
{
    std::unique_lock<std::mutex> g(chLock_);
    modBodyCh_.push_back({id, name, cls, mods});
    chCv_.wait(g, [&] { return !modBodyResultCh_.empty(); });
    const ModBodyResult r = modBodyResultCh_.front();
    modBodyResultCh_.pop_front();
    return r;
}

bool BodyCollection::HandleModBody()
{
    ModReq req;
    {
        std::lock_guard<std::mutex> g(chLock_);
        if (modBodyCh_.empty()) return false;
        req = modBodyCh_.front();
        modBodyCh_.pop_front();
    }
    if (syncFromDevice) syncFromDevice();
    int found = 0, modified = 0;
    for (auto &b : arr_) {
        if ((!req.cls.empty() && req.cls == b->Class) || (!req.name.empty() && req.name == b->Name) ||
            req.id == b->Id) {
            found++;
            if (b->ApplyMods(req.mods)) modified++;
        }
    }
    ModBodyResult r;
    if (found == 0) r = ModBodyResult::NoMatch;
    else if (modified == 0) r = ModBodyResult::ModNone;
    else if (found == modified) r = ModBodyResult::ModAll;
    else r = ModBodyResult::ModSome;
    std::lock_guard<std::mutex> g(chLock_);
    modBodyResultCh_.push_back(r);
    chCv_.notify_all();
    return true;
}

// ---------------------------------------------------------------- ResultQueueHolder
ResultQueueHolder::ResultQueueHolder(int maxQueues) : maxQueues_(maxQueues), physCap_(maxQueues + 1) {}

std::pair<ResultQueuePtr, bool> ResultQueueHolder::NewResultQueue()
{
    std::lock_guard<std::mutex> g(lock_);
    if ((int)ch_.size() >= maxQueues_) return {nullptr, false};
    auto q = std::make_shared<ResultQueue>();
    q->QueueNum = queueNum_++;
    return {q, true};
}

void ResultQueueHolder::Add(ResultQueuePtr q)
{
    std::lock_guard<std::mutex> g(lock_);
    if ((int)ch_.size() >= physCap_) {
        std::fprintf(stderr, "No queue capacity len=%zu cap=%d max=%d\n", ch_.size(), physCap_, maxQueues_);
        std::abort();  // log.Fatalf in the reference (resultqueue.go:62-65)
    }
    ch_.push_back(std::move(q));
}

std::pair<ResultQueuePtr, bool> ResultQueueHolder::Next()
{
    std::lock_guard<std::mutex> g(lock_);
    if (ch_.empty()) return {nullptr, false};
    auto q = ch_.front();
    ch_.pop_front();
    return {q, true};
}

int ResultQueueHolder::MaxQueues()
{
    std::lock_guard<std::mutex> g(lock_);
    return maxQueues_;
}

int ResultQueueHolder::Len()
{
    std::lock_guard<std::mutex> g(lock_);
    return (int)ch_.size();
}

bool ResultQueueHolder::Resize(int maxQueues)
{
    std::lock_guard<std::mutex> g(lock_);
    if (maxQueues == maxQueues_) return false;
    // Physical room always covers the current content plus the one queue a concurrent
    // NewResultQueue may already have granted (the cycle in flight).  The reference only adds that
    // slack when shrinking *below* the content (resultqueue.go:176-182); shrinking to exactly the
    // content (maxQueues == curLen) leaves none and its Add then dies with "No queue capacity".
    physCap_ = std::max(maxQueues, (int)ch_.size()) + 1;
    maxQueues_ = maxQueues;
    return true;
}

}  // namespace nbodygo
