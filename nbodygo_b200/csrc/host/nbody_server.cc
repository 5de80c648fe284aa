// nbody_server.cc — headless twin of cmd/server/main.go for the GPU path.
//
// Accepts the reference's compute-related flags in both `--opt v` and `--opt=v` forms
// (cmd/server/main.go:113-236): --sim-name/-n, --sim-args/-a, --collision/-c, --bodies/-b,
// --threads/-t (accepted, no-op), --scaling/-m, --csv/-f, --run-millis/-u, --no-render/-r
// (always on: there is no renderer here), --no-barnes-hut (always brute force).
// Additions: --gpu=<device>, --seed=<n>, --iterations=<n>, --dump-csv=<path> (initial bodies),
// --dump-final-csv=<path> (surviving bodies after the run in the reference's CSV format),
// --dump-final-state=<path> / --resume-state=<path> (the sidecar with what that CSV cannot carry: together a
// bit-exact checkpoint, the reference has none), --restitution=<R> (the runner's R; the reference sets it over
// gRPC only).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "nbody_host.h"

using namespace nbodygo;

int main(int argc, char **argv)
{
    std::string simName = "Sim1", simArgs, csvPath, dumpCsv, dumpFinalCsv, dumpFinalState, resumeState;
    double restitution = 1;
    bool haveR = false;
    CollisionBehavior behavior = Elastic;
    BodyColor color = Random;
    int bodyCount = 1000, runMillis = -1, iterations = 0, device = 0;
    double scaling = .000000001;
    uint64_t seed = 1;

    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], v;
        const size_t eq = a.find('=');
        bool hasVal = false;
        if (eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); hasVal = true; }
        auto val = [&]() -> std::string {
            if (hasVal) return v;
            if (i + 1 < argc) return argv[++i];
            std::fprintf(stderr, "missing value for %s\n", a.c_str());
            std::exit(2);
        };
        if (a == "--sim-name" || a == "-n") simName = val();
        else if (a == "--sim-args" || a == "-a") simArgs = val();
        else if (a == "--collision" || a == "-c") behavior = ParseCollisionBehavior(val());
        else if (a == "--bodies" || a == "-b") bodyCount = std::atoi(val().c_str());
        else if (a == "--threads" || a == "-t") (void)val();
        else if (a == "--scaling" || a == "-m") scaling = (double)(float)std::atof(val().c_str());  // float32, main.go:199
        else if (a == "--csv" || a == "-f") csvPath = val();
        else if (a == "--body-color" || a == "-l") color = ParseBodyColor(val());
        else if (a == "--run-millis" || a == "-u") runMillis = std::atoi(val().c_str());
        else if (a == "--no-render" || a == "-r" || a == "--no-barnes-hut") {}
        else if (a == "--gpu") device = std::atoi(val().c_str());
        else if (a == "--seed") seed = std::strtoull(val().c_str(), nullptr, 10);
        else if (a == "--iterations") iterations = std::atoi(val().c_str());
        else if (a == "--dump-csv") dumpCsv = val();
        else if (a == "--dump-final-csv") dumpFinalCsv = val();
        else if (a == "--dump-final-state") dumpFinalState = val();
        else if (a == "--resume-state") resumeState = val();
        else if (a == "--restitution") { restitution = std::atof(val().c_str()); haveR = true; }
        else {
            std::fprintf(stderr, "unknown option %s\n", a.c_str());
            return 2;
        }
    }

    std::vector<BodyPtr> bodies;
    if (!csvPath.empty()) {
        bodies = FromCsv(csvPath, bodyCount, behavior, color);
    } else if (simName != "empty") {
        std::string nm = simName;
        if (!nm.empty()) nm[0] = (char)std::toupper((unsigned char)nm[0]);
        bodies = Generate(nm, bodyCount, behavior, color, simArgs, seed);
        if (bodies.empty()) {
            std::fprintf(stderr, "ERROR: could not build sim specified on the command line: %s\n", simName.c_str());
            return 1;
        }
    }
    if (!dumpCsv.empty()) WriteCsv(dumpCsv, bodies);
    RunState start;
    bool haveStart = false;
    if (!resumeState.empty()) {
        if (!ReadState(resumeState, bodies, start)) {
            std::fprintf(stderr, "ERROR: %s does not match the bodies read from the CSV\n", resumeState.c_str());
            return 1;
        }
        haveStart = true;
    }
    if (haveR) {  // overrides the R of a resumed run, like a set-restitution-coefficient call before its first cycle
        start.R = restitution;
        haveStart = true;
    }
    try {
        const HeadlessResult r = RunHeadless(bodies, scaling, runMillis, iterations, device, false,
                                             haveStart ? &start : nullptr);
        std::printf("bodies: %zu -> %d\ninteractions/s: %.6e\n", bodies.size(), r.finalBodies, r.interactionsPerSec);
        if (!dumpFinalCsv.empty()) {
            std::vector<BodyPtr> alive;
            for (auto &b : r.bodies)
                if (b->Exists) alive.push_back(b);
            WriteCsv(dumpFinalCsv, alive);
        }
        if (!dumpFinalState.empty()) {
            std::vector<BodyPtr> alive;
            for (auto &b : r.bodies)
                if (b->Exists) alive.push_back(b);
            WriteState(dumpFinalState, alive, r.state);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
