#!/bin/bash
# Frame rates of the reference-sized sims through the full C++ host loop (nbody_server =
# ComputationRunner -> GpuStepper -> nb_step, result queues drained, Cycle every frame).
#   tools/sim_rates.sh [iterations]
IT=${1:-3000}
S=nbodygo_b200/bin/nbody_server
run() {  # name bodies collision
    echo "== $1 bodies=$2 collision=$3"
    $S --sim-name "$1" --bodies="$2" --collision="$3" --no-render --no-barnes-hut --iterations="$IT" --seed=7 2>/dev/null |
        grep -E "computations:|device ms|frames per second|uploads|^bodies"
}
run Sim3 1001 elastic
run Sim1 1001 elastic
run Sim1 3001 elastic
run Sim2 5000 elastic   # the sun swallows what passes through it
run Sim2 5000 subsume   # every overlap merges: the cluster collapses into a few bodies
run Sim5 6 fragment
