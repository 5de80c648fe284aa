package sim

// parity_dump_test.go — runs the UNMODIFIED reference path (Body.Compute → ProcessMods → Body.Update →
// Cycle, brute force: bc.Tree stays nil as with --no-barnes-hut) on a CSV the B200 repo wrote
// (tools/write_inputs.py) and dumps every bit of the result, so that tools/compare_go_dump.py can pin the
// CPU oracle and the CUDA path against the real thing.  Needs cmd/body/gpu_accessors.go (exported
// accessors only).  Single worker, array order: the deterministic idealisation of the goroutine pool
// (same arrival order as one slice, cmd/runner/workpool.go:103-110).
//
//   NB_DUMP_CSV=c3_2000.csv NB_DUMP_CYCLES=3 NB_DUMP_TS=1e-9 NB_DUMP_R=1 NB_DUMP_OUT=go_dump.txt \
//       go test ./cmd/sim -run TestParityDump -count=1
//
// Records (one per line; floats as %016x of math.Float64bits):
//   H n ts R cycles          header
//   C cycle n                start of a cycle
//   E kind a b               deferred events in the order ProcessMods handles them (Front→Next); kind 0
//                            collision, 1 subsume; a, b = array indices of b1, b2
//   F i fx fy fz             Body.fx,fy,fz after Compute
//   S i x y z vx vy vz mass exists     after Update
//   N n                      body count after Cycle
//
// The event channel of the reference holds 1000 events and DROPS on full (body_collection.go:82-88): the
// test waits for the channel to drain after every Compute, so a drop can only happen when a single body
// raises more than 1000 events; the reference then logs "[ERROR] Attempt to enqueue event on full event
// channel" — a dump taken with that line in the log is not lossless.

import (
	"bufio"
	"fmt"
	"math"
	"nbodygo/cmd/body"
	"nbodygo/cmd/globals"
	"os"
	"strconv"
	"testing"
	"time"
)

func envFloat(name string, dflt float64) float64 {
	if v, err := strconv.ParseFloat(os.Getenv(name), 64); err == nil {
		return v
	}
	return dflt
}

func hex(f float64) string { return fmt.Sprintf("%016x", math.Float64bits(f)) }

// waits until everything Enqueue sent has reached the deferred-event list
func drainEvents(bc *body.BodyCollection) {
	for bc.EventBacklog() > 0 {
		time.Sleep(50 * time.Microsecond)
	}
	// handleEvents may hold the last event between the channel and the list: wait for a stable length
	last, stable := -1, 0
	for stable < 5 {
		n := len(bc.PendingEvents())
		if n == last {
			stable++
		} else {
			last, stable = n, 0
		}
		time.Sleep(200 * time.Microsecond)
	}
}

func TestParityDump(t *testing.T) {
	csvPath := os.Getenv("NB_DUMP_CSV")
	if csvPath == "" {
		t.Skip("NB_DUMP_CSV not set")
	}
	cycles := int(envFloat("NB_DUMP_CYCLES", 1))
	ts := envFloat("NB_DUMP_TS", 1e-9)
	R := envFloat("NB_DUMP_R", 1)
	outPath := os.Getenv("NB_DUMP_OUT")
	if outPath == "" {
		outPath = "go_dump.txt"
	}
	bodies := FromCsv(csvPath, math.MaxInt32, globals.Elastic, globals.Random)
	if bodies == nil {
		t.Fatalf("could not read %s", csvPath)
	}
	f, err := os.Create(outPath)
	if err != nil {
		t.Fatal(err)
	}
	defer f.Close()
	w := bufio.NewWriter(f)
	defer w.Flush()

	bc := body.NewSimBodyCollection(bodies)
	fmt.Fprintf(w, "H %d %s %s %d\n", len(bodies), hex(ts), hex(R), cycles)
	for c := 0; c < cycles; c++ {
		arr := bc.GetArray()
		fmt.Fprintf(w, "C %d %d\n", c, len(arr))
		// Body.Compute for every body, array order, one worker (computation-runner.go:297-311)
		for _, b := range arr {
			b.Compute(bc)
			drainEvents(bc)
		}
		for _, e := range bc.PendingEvents() {
			if e.Kind == 0 || e.Kind == 1 {
				fmt.Fprintf(w, "E %d %d %d\n", e.Kind, e.A, e.B)
			}
		}
		for i, b := range arr {
			fx, fy, fz := b.Forces()
			fmt.Fprintf(w, "F %d %s %s %s\n", i, hex(fx), hex(fy), hex(fz))
		}
		bc.ProcessMods() // computation-runner.go:316
		for i, b := range arr {
			b.Update(ts, R) // computation-runner.go:317-320
			ex := 0
			if b.Exists {
				ex = 1
			}
			fmt.Fprintf(w, "S %d %s %s %s %s %s %s %s %d\n", i, hex(b.X), hex(b.Y), hex(b.Z), hex(b.Vx), hex(b.Vy),
				hex(b.Vz), hex(b.Mass), ex)
		}
		bc.Cycle(R) // computation-runner.go:322
		fmt.Fprintf(w, "N %d\n", bc.Count())
	}
}
