// host_sim.cc — CSV channel, seeded sim generators and the headless run loop (cmd/sim).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <random>
#include <sstream>

#include "nbody_host.h"

namespace nbodygo {

static const double solarMass = 1.98892e30;  // simgen.go:34-36

static std::string trim(const std::string &s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

static bool parseFloatStrict(const std::string &s, double &out)
{
    if (s.empty()) return false;
    char *end = nullptr;
    out = std::strtod(s.c_str(), &end);
    return end != s.c_str() && *end == '\0';
}

// strconv.ParseBool: 1,t,T,TRUE,true,True / 0,f,F,FALSE,false,False
static bool parseBoolStrict(const std::string &s, bool &out)
{
    if (s == "1" || s == "t" || s == "T" || s == "TRUE" || s == "true" || s == "True") { out = true; return true; }
    if (s == "0" || s == "f" || s == "F" || s == "FALSE" || s == "false" || s == "False") { out = false; return true; }
    return false;
}

std::vector<BodyPtr> FromCsv(const std::string &csvPath, int bodyCount, CollisionBehavior defaultBehavior,
                             BodyColor defaultColor)
{
    std::vector<BodyPtr> bodies;
    std::ifstream f(csvPath);
    if (!f) {
        std::fprintf(stderr, "Error opening csv: %s\n", csvPath.c_str());
        return bodies;
    }
    std::string line;
    int lines = 0;
    while (lines < bodyCount && std::getline(f, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> fields;
        std::stringstream ss(line);
        std::string cell;
        while (std::getline(ss, cell, ',')) fields.push_back(trim(cell));
        if (!line.empty() && line.back() == ',') fields.push_back("");
        double v[8];
        bool ok = fields.size() >= 8;
        for (int k = 0; ok && k < 8; ++k) ok = parseFloatStrict(fields[k], v[k]);
        bool isSun = false;
        if (ok && fields.size() >= 9) ok = parseBoolStrict(fields[8], isSun);
        CollisionBehavior beh = defaultBehavior;
        if (ok && fields.size() >= 10) beh = ParseCollisionBehavior(fields[9]);
        BodyColor color = defaultColor;
        if (ok && fields.size() >= 11) color = ParseBodyColor(fields[10]);
        double fragFactor = 0, fragStep = 0;
        if (ok && fields.size() >= 12) ok = parseFloatStrict(fields[11], fragFactor);
        if (ok && fields.size() >= 13) ok = parseFloatStrict(fields[12], fragStep);
        if (!ok) continue;  // the reference recovers from the parse panic and skips the record
        auto b = NewBody(NextId(), v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], beh, color, fragFactor, fragStep,
                         false, "", "", false);
        if (isSun) b->SetSun(100);
        bodies.push_back(b);
        lines++;
    }
    return bodies;
}

bool WriteCsv(const std::string &csvPath, const std::vector<BodyPtr> &bodies)
{
    static const char *behNames[] = {"none", "subsume", "elastic", "fragment"};
    static const char *colNames[] = {"random", "black", "white", "darkgray", "gray", "lightgray", "red", "green",
                                     "blue", "yellow", "magenta", "cyan", "orange", "brown", "pink"};
    FILE *f = std::fopen(csvPath.c_str(), "w");
    if (!f) return false;
    std::fprintf(f, "# x,y,z,vx,vy,vz,mass,radius,is_sun,collision_behavior,color,frag_factor,frag_step\n");
    for (auto &b : bodies)
        std::fprintf(f, "%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%s,%s,%s,%.17g,%.17g\n", b->X, b->Y, b->Z,
                     b->Vx, b->Vy, b->Vz, b->Mass, b->Radius, b->IsSun ? "true" : "false", behNames[b->Behavior],
                     colNames[b->Color], b->FragFactor, b->FragStep);
    std::fclose(f);
    return true;
}

// ---------------------------------------------------------------- state sidecar (see RunState in nbody_host.h)
bool WriteState(const std::string &path, const std::vector<BodyPtr> &bodies, const RunState &rs)
{
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    std::fprintf(f, "#nbstate v1\t%zu\t%d\t%d\t%a\n", bodies.size(), rs.nextId, rs.cycle, rs.R);
    for (auto &b : bodies)
        std::fprintf(f, "%d\t%s\t%s\t%d\t%d\t%a\t%d\t%a\t%a\t%a\t%d\t%a\t%a\t%a\t%a\t%a\t%a\t%a\n", b->Id,
                     b->Name.c_str(), b->Class.c_str(), b->Pinned ? 1 : 0, b->WithTelemetry ? 1 : 0, b->r,
                     b->fragmenting ? 1 : 0, b->fragInfo.radius, b->fragInfo.newRadius, b->fragInfo.mass,
                     b->fragInfo.fragments, b->fragInfo.x, b->fragInfo.y, b->fragInfo.z, b->fx, b->fy, b->fz,
                     b->intensity);
    std::fclose(f);
    return true;
}

bool ReadState(const std::string &path, std::vector<BodyPtr> &bodies, RunState &rs)
{
    std::ifstream f(path);
    if (!f) return false;
    auto split = [](const std::string &line) {
        std::vector<std::string> out;
        size_t a = 0;
        for (;;) {
            const size_t t = line.find('\t', a);
            out.push_back(line.substr(a, t == std::string::npos ? std::string::npos : t - a));
            if (t == std::string::npos) break;
            a = t + 1;
        }
        return out;
    };
    std::string line;
    if (!std::getline(f, line)) return false;
    const auto h = split(line);
    if (h.size() != 5 || h[0] != "#nbstate v1" || (size_t)std::strtoull(h[1].c_str(), nullptr, 10) != bodies.size())
        return false;
    rs.nextId = std::atoi(h[2].c_str());
    rs.cycle = std::atoi(h[3].c_str());
    rs.R = std::strtod(h[4].c_str(), nullptr);
    for (auto &b : bodies) {
        if (!std::getline(f, line)) return false;
        const auto c = split(line);
        if (c.size() != 18) return false;
        auto d = [&](int k) { return std::strtod(c[(size_t)k].c_str(), nullptr); };  // strtod reads C99 hex floats
        b->Id = std::atoi(c[0].c_str());
        b->Name = c[1];
        b->Class = c[2];
        b->Pinned = c[3] == "1";
        b->WithTelemetry = c[4] == "1";
        b->r = d(5);
        b->fragmenting = c[6] == "1";
        b->fragInfo = FragInfo{d(7), d(8), d(9), std::atoi(c[10].c_str()), d(11), d(12), d(13)};
        b->fx = d(14); b->fy = d(15); b->fz = d(16);
        b->intensity = d(17);
    }
    return true;
}

// ---------------------------------------------------------------- generators (simgen.go), seeded
namespace {
struct Rng {
    std::mt19937_64 g;
    std::uniform_real_distribution<double> u{0.0, 1.0};
    explicit Rng(uint64_t seed) : g(seed) {}
    double f() { return u(g); }
    void even(double cx, double cy, double cz, double radius, double out[3])
    {  // util.GetVectorEven
        double x, y, z, d = 2;
        while (d > 1) {
            x = f() * 2 - 1; y = f() * 2 - 1; z = f() * 2 - 1;
            d = x * x + y * y + z * z;
        }
        out[0] = x * radius + cx; out[1] = y * radius + cy; out[2] = z * radius + cz;
    }
};

void addSun(std::vector<BodyPtr> &bodies, double x, double y, double z, double mass, double radius, double intensity)
{  // createSunAndAddToList, simgen.go:424-429
    auto b = NewBody(NextId(), x, y, z, -3, -3, -5, mass, radius, Subsume, White, 0, 0, false, "the-sun", "", true);
    b->SetSun(intensity);
    bodies.push_back(b);
}

std::vector<std::string> splitArgs(const std::string &s)
{
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string c;
    while (std::getline(ss, c, ',')) out.push_back(c);
    return out;
}
}  // namespace

std::vector<BodyPtr> Generate(const std::string &simName, int bodyCount, CollisionBehavior behavior, BodyColor color,
                              const std::string &simArgs, uint64_t seed)
{
    Rng rng(seed);
    std::vector<BodyPtr> bodies;
    const auto args = splitArgs(simArgs);
    if (simName == "Sim1") {  // simgen.go:93-168: four clumps around a sun
        double clumpRadius = 30, dist = 200;
        if (args.size() > 0) clumpRadius = SafeParseFloat(args[0], clumpRadius);
        if (args.size() > 1) dist = SafeParseFloat(args[1], dist);
        const double V = 958000000;
        for (int i = -1; i <= 1; i += 2)
            for (int j = -1; j <= 1; j += 2) {
                const double xc = dist * i, zc = dist * j;
                double vx, vz, y;
                BodyColor c = color;
                if (i == -1 && j == -1) { vx = -V; vz = V; y = 100; if (color == Random) c = Red; }
                else if (i == -1 && j == 1) { vx = V; vz = V; y = -100; if (color == Random) c = Yellow; }
                else if (i == 1 && j == 1) { vx = V; vz = -V; y = 100; if (color == Random) c = Lightgray; }
                else { vx = -V; vz = -V; y = -100; if (color == Random) c = Cyan; }
                for (int k = 0; k < bodyCount / 4; ++k) {
                    const double vy = .5 - rng.f();
                    const double f = rng.f();
                    const double radius = (double)k < (double)bodyCount * .0025 ? 8 * f : 3 * f;
                    const double mass = radius * solarMass * .000005;
                    double v[3];
                    rng.even(xc, y, zc, clumpRadius, v);
                    bodies.push_back(NewBody(NextId(), v[0], v[1], v[2], vx, vy, vz, mass, radius, behavior, c, 0, 0,
                                             false, "", "", false));
                }
            }
        addSun(bodies, 0, 0, 0, 25 * solarMass * .11, 35, 100);
    } else if (simName == "Sim2") {  // :180-195: sun + one cluster on a close pass
        addSun(bodies, 0, 0, 0, 25 * solarMass * .1, 25, 100);
        for (int i = 1; i < bodyCount; ++i) {
            double v[3];
            rng.even(500, 500, 500, 50, v);
            const double mass = rng.f() * solarMass * .000005;
            const double radius = rng.f() * 4;
            bodies.push_back(NewBody(NextId(), v[0], v[1], v[2], -1124500000, -824500000, -1124500000, mass, radius,
                                     behavior, color, 1, 1, false, "", "", false));
        }
    } else if (simName == "Sim3") {  // :218-254: far sun + two colliding clusters
        double radius = 50, mass = 90E25;
        if (args.size() > 0) radius = SafeParseFloat(args[0], radius);
        if (args.size() > 1) mass = SafeParseFloat(args[1], mass);
        addSun(bodies, 100000, 100000, 100000, 1, 500, 4E5);
        for (int j = -1; j <= 1; j += 2)
            for (int i = 0; i < bodyCount / 2; ++i) {
                BodyColor c = color == Random ? (j == 1 ? Yellow : Red) : color;
                double v[3];
                rng.even(j * 70.0, j * 70.0, j * 70.0, radius, v);
                bodies.push_back(NewBody(NextId(), v[0], v[1], v[2], j * 121185000.0, j * 121185000.0,
                                         j * -121185000.0, mass, 5, behavior, c, 1, 1, false, "", "", false));
            }
    } else if (simName == "Sim4") {  // :289-303: a line of bodies past a sun
        addSun(bodies, 0, 0, 0, solarMass, 30, 90);
        for (int i = 1; i < bodyCount; ++i)
            bodies.push_back(NewBody(NextId(), (double)(i * 4) + 100, 0, 0, 0, 0, -824500000 + (double)(i * 1E6), 9e5,
                                     2, behavior, color, 1, 1, false, "", "", false));
    } else if (simName == "Sim5") {  // :322-374: planet, moons, fragmenting impactor
        double fragFactor = .01, fragStep = 1000;
        if (args.size() >= 1) fragFactor = SafeParseFloat(args[0], fragFactor);
        if (args.size() >= 2) fragStep = SafeParseFloat(args[1], fragStep);
        addSun(bodies, 100000, 100000, 1000, 1, 500, 4E5);
        bodies.push_back(NewBody(NextId(), 0, 0, 0, 12, 12, 12, 9E30, 145, Elastic, Red, 0, 0, false, "", "", false));
        bodies.push_back(NewBody(NextId(), 50, 0, -420, -980000000, 12, -500000000, 9E20, 35, Subsume, Lightgray, 0,
                                 0, false, "", "", false));
        bodies.push_back(NewBody(NextId(), -400, 50, 405, 530000000, -313000000, 520000000, 9E19, 5, Elastic, Blue, 0,
                                 0, false, "", "", false));
        bodies.push_back(NewBody(NextId(), 70, 0, -520, -880000000, -10000, -300000000, 11E22, 15, Elastic, Green, 0,
                                 0, false, "", "", false));
        bodies.push_back(NewBody(NextId(), 900, -900, 900, -450000000, 723000000, -350000000, 9E12, 10, Fragment,
                                 Yellow, fragFactor, fragStep, false, "", "", false));
    } else if (simName == "SimTest") {  // :381-404
        addSun(bodies, 20000, 20000, 20000, 1, 500, 10000);
        bodies.push_back(NewBody(NextId(), 0, 0, 0, 0, 0, 0, 9E29, 60, behavior, Red, 0, 0, false, "", "", false));
        bodies.push_back(NewBody(NextId(), -350, 350, 0, 530000000, -500000000, 0, 9E29, 60, behavior, Green, 0, 0,
                                 false, "", "", false));
        bodies.push_back(NewBody(NextId(), 350, 350, 0, -530000000, -500000000, 0, 9E29, 60, behavior, Yellow, 0, 0,
                                 false, "", "", false));
    }
    return bodies;  // unknown name → empty, like Generate's nil (simgen.go:62-64)
}

// ---------------------------------------------------------------- headless run (nbodysim.go:78-133)
HeadlessResult RunHeadless(std::vector<BodyPtr> bodies, double timeScaling, int runMillis, int maxIterations,
                           int device, bool quiet, const RunState *start)
{
    HeadlessResult out;
    BodyCollection bc(bodies);
    ResultQueueHolder rqh(10);
    ComputationRunner runner(1, timeScaling, false, &rqh, &bc, device);
    if (start) {  // continue a dumped run: cycle counter, id generator and R are part of its state
        if (start->cycle >= 0) bc.setCycle(start->cycle);
        if (start->nextId >= 0) SetNextId(start->nextId);
        runner.SetCoefficientOfRestitution(start->R);  // picked up at the top of the first cycle, like the gRPC setter
    }
    if (maxIterations > 0) runner.SetMaxIterations(maxIterations);
    const auto t0 = std::chrono::steady_clock::now();
    runner.Start();
    const double n0 = (double)bodies.size();
    while (runner.Running()) {
        // waitForSimEnd: with rendering off somebody must drain the queues
        while (rqh.Next().second) {}
        std::this_thread::sleep_for(std::chrono::microseconds(200));
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (runMillis > 0 && ms >= runMillis) break;
    }
    runner.Stop();
    while (rqh.Next().second) {}
    out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out.computations = runner.Computations();
    out.iterations = runner.Iterations();
    out.fps = out.seconds > 0 ? out.computations / out.seconds : 0;
    out.interactionsPerSec = out.fps * n0 * (n0 - 1);
    out.finalBodies = bc.Count();
    out.bodies = bc.GetArray();
    out.state.cycle = bc.cycle();
    out.state.R = runner.CoefficientOfRestitution();
    out.state.nextId = NextId();  // consumes one id: only the state of a finished run is read
    if (!quiet) runner.PrintStats();
    return out;
}

}  // namespace nbodygo
