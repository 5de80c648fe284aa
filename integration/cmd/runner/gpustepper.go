//go:build gpu

package runner

// gpustepper.go — cgo binding of libnbody_b200.so (include/nbody_b200.h): the B200 replacement for the
// goroutine work pool.  One Step() call stands in for computation-runner.go:285-320 (partition →
// submitSlice → wait → ProcessMods → Update loop).  The device image of BodyCollection.arr stays resident
// between cycles; only the float32 Renderable snapshot (13 B/body) crosses the host link every cycle, and
// adds / deletes / mods are patched in (nb_append / nb_compact / nb_patch) instead of re-uploading.
//
// Build:  CGO_CFLAGS="-I<repo>/include" CGO_LDFLAGS="-L<repo>/nbodygo_b200 -lnbody_b200 \
//         -Wl,-rpath,<repo>/nbodygo_b200" go build -tags gpu ./cmd/server
//
// The C++ twin of this file, compiled and tested on B200, is nbodygo_b200/csrc/host/host_runner.cc
// (GpuStepper::Step / AfterCycle / Reserve).

/*
#cgo LDFLAGS: -lnbody_b200
#include <stdlib.h>
#include "nbody_b200.h"
*/
import "C"

import (
	"fmt"
	"log"
	"nbodygo/cmd/body"
	"runtime"
	"unsafe"
)

// GpuStepper owns one device image of BodyCollection.arr.  All calls come from the runner goroutine, which
// holds runtime.LockOSThread() (the handle is not thread-safe and CUDA contexts are per thread).
type GpuStepper struct {
	h       C.nb_handle
	device  int
	n       int   // bodies on the device
	cap     int   // device capacity
	pairCap int64 // event capacity of the handle
	// host staging (SoA); reused every cycle, never retained by the library
	x, y, z, vx, vy, vz, mass, radius, rest, ff, fs []float64
	fx, fy, fz                                      []float64
	beh, flags                                      []uint8
	// library-owned pinned snapshot buffers that every nb_step fills (nb_render_buffers)
	pinXyz    *C.float
	pinExists *C.uint8_t
	dirty     bool // the Go bodies changed (mods / deletes / errors): re-upload before the next step
	hostStale bool // the device advanced since the Go bodies were last refreshed
	Failed    uint // device steps that failed (nothing published, no Cycle)
}

// NewGpuStepper takes NewWorkPool's place (workpool.go:129-142).  There is no CPU fallback: without a usable
// device the error is returned and the caller decides (cmd/server exits).
func NewGpuStepper(device, capacity int) (*GpuStepper, error) {
	runtime.LockOSThread()
	g := &GpuStepper{device: device, dirty: true}
	if err := g.recreate(capacity, 0); err != nil {
		return nil, err
	}
	return g, nil
}

func (g *GpuStepper) recreate(capacity int, pairCapacity int64) error {
	if g.h != nil {
		C.nb_destroy(g.h)
		g.h = nil
	}
	if rc := C.nb_create(C.int(g.device), C.int64_t(capacity), C.int64_t(pairCapacity), &g.h); rc != C.NB_OK {
		return fmt.Errorf("nb_create: %s", C.GoString(C.nb_last_error(nil)))
	}
	g.cap = capacity
	g.pairCap = pairCapacity
	if g.pairCap <= 0 {
		g.pairCap = 4*int64(capacity) + 65536 // nb_create's default
	}
	g.n = 0
	g.dirty = true
	if rc := C.nb_render_buffers(g.h, &g.pinXyz, &g.pinExists); rc != C.NB_OK {
		g.pinXyz, g.pinExists = nil, nil
	}
	return nil
}

// Close releases the device image.
func (g *GpuStepper) Close() {
	if g.h != nil {
		C.nb_destroy(g.h)
		g.h = nil
	}
}

// MarkDirty: HandleModBody / processDeletes changed Go bodies — re-upload before the next step.
func (g *GpuStepper) MarkDirty() { g.dirty = true }

func (g *GpuStepper) lastError() string { return C.GoString(C.nb_last_error(g.h)) }

func (g *GpuStepper) grow(n int) {
	f64 := []*[]float64{&g.x, &g.y, &g.z, &g.vx, &g.vy, &g.vz, &g.mass, &g.radius, &g.rest, &g.ff, &g.fs,
		&g.fx, &g.fy, &g.fz}
	if cap(g.x) >= n {
		for _, s := range f64 {
			*s = (*s)[:n]
		}
		g.beh, g.flags = g.beh[:n], g.flags[:n]
		return
	}
	c := n + n/2 + 16
	for _, s := range f64 {
		*s = make([]float64, n, c)
	}
	g.beh, g.flags = make([]uint8, n, c), make([]uint8, n, c)
}

func dptr(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&s[0]))
}

func bptr(s []uint8) *C.uint8_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&s[0]))
}

func flagsOf(b *body.Body) uint8 {
	var f uint8
	if b.Exists {
		f |= C.NB_F_EXISTS
	}
	if b.IsFragmenting() {
		f |= C.NB_F_FRAGMENTING
	}
	if b.Pinned {
		f |= C.NB_F_PINNED
	}
	if b.IsSun {
		f |= C.NB_F_SUN
	}
	if b.WithTelemetry {
		f |= C.NB_F_TELEMETRY
	}
	return f
}

func (g *GpuStepper) stage(i int, b *body.Body) {
	g.x[i], g.y[i], g.z[i] = b.X, b.Y, b.Z
	g.vx[i], g.vy[i], g.vz[i] = b.Vx, b.Vy, b.Vz
	g.mass[i], g.radius[i] = b.Mass, b.Radius
	g.rest[i], g.ff[i], g.fs[i] = b.Restitution(), b.FragFactor, b.FragStep
	g.beh[i], g.flags[i] = uint8(b.CollisionBehavior), flagsOf(b)
}

// upload marshals []*Body → SoA → device (NewSimBodyCollection's place).  The reference's array simply
// grows (body_collection.go:273-291); so does the device image.
func (g *GpuStepper) upload(arr []*body.Body) {
	n := len(arr)
	if n > g.cap {
		log.Printf("[INFO] device capacity %d -> %d bodies\n", g.cap, 2*n+4096)
		if err := g.recreate(2*n+4096, 0); err != nil {
			log.Fatalf("[ERROR] %v", err)
		}
	}
	g.grow(n)
	anyFrag := false
	for i, b := range arr {
		g.stage(i, b)
		g.fx[i], g.fy[i], g.fz[i] = b.Forces()
		anyFrag = anyFrag || b.IsFragmenting()
	}
	if rc := C.nb_upload(g.h, C.int64_t(n), dptr(g.x), dptr(g.y), dptr(g.z), dptr(g.vx), dptr(g.vy), dptr(g.vz),
		dptr(g.mass), dptr(g.radius), dptr(g.rest), dptr(g.ff), dptr(g.fs), bptr(g.beh), bptr(g.flags)); rc != C.NB_OK {
		log.Printf("[ERROR] nb_upload: %s", g.lastError())
	} else if anyFrag {
		// nb_upload starts every body at fx = fy = fz = 0; a fragmenting body keeps applying the force of
		// its last Compute (body.go:152-155)
		if rc := C.nb_set_forces(g.h, 0, C.int64_t(n), dptr(g.fx), dptr(g.fy), dptr(g.fz)); rc != C.NB_OK {
			log.Printf("[ERROR] nb_set_forces: %s", g.lastError())
		}
	}
	g.n, g.dirty, g.hostStale = n, false, false
}

// SyncToHost refreshes the Go bodies from the device (GetBody, mods, deletes, end of run): the device owns
// X..Vz, r and fx..fz between syncs.
func (g *GpuStepper) SyncToHost(bc *body.BodyCollection) {
	if !g.hostStale {
		return
	}
	arr := bc.GetArray()
	n := len(arr)
	if g.n < n {
		n = g.n
	}
	if n == 0 {
		g.hostStale = false
		return
	}
	g.grow(g.n)
	if rc := C.nb_download_state(g.h, dptr(g.x), dptr(g.y), dptr(g.z), dptr(g.vx), dptr(g.vy), dptr(g.vz), nil, nil,
		dptr(g.rest), nil, nil); rc != C.NB_OK {
		log.Printf("[ERROR] nb_download_state: %s", g.lastError())
		return
	}
	haveF := C.nb_get_forces(g.h, dptr(g.fx), dptr(g.fy), dptr(g.fz)) == C.NB_OK
	for i := 0; i < n; i++ {
		b := arr[i]
		b.X, b.Y, b.Z, b.Vx, b.Vy, b.Vz = g.x[i], g.y[i], g.z[i], g.vx[i], g.vy[i], g.vz[i]
		b.SetRestitution(g.rest[i])
		b.ClearCollided()
		if haveF {
			b.SetForces(g.fx[i], g.fy[i], g.fz[i])
		}
	}
	g.hostStale = false
}

// Step replaces computation-runner.go:285-320.  It returns false (after logging [ERROR]) when the device
// step failed: nothing was published into rq and the caller must not Cycle.
func (g *GpuStepper) Step(bc *body.BodyCollection, timeScaling, R float64, rq *ResultQueue) bool {
	arr := bc.GetArray()
	inStep := !g.dirty && len(arr) == g.n
	// fragment() copies the body's CURRENT velocity into its fragments (fragcalc.go:97)
	if g.hostStale && inStep {
		for _, b := range arr {
			if b.Exists && b.IsFragmenting() {
				g.SyncToHost(bc)
				break
			}
		}
	}
	// fragmenting bodies spawn their fragments on the host, as Body.Compute does (body.go:152-155)
	for i, b := range arr {
		if b.Exists && b.IsFragmenting() {
			b.Fragment(bc)
			if !b.Exists && inStep { // fully fragmented (fragcalc.go:114-116): one flag byte to the device
				fl := flagsOf(b)
				if rc := C.nb_patch(g.h, C.int64_t(i), 1, nil, nil, nil, nil, nil, nil, nil, nil, nil, nil, nil, nil,
					(*C.uint8_t)(unsafe.Pointer(&fl))); rc != C.NB_OK {
					g.dirty = true
				}
			}
		}
	}
	if g.dirty || len(arr) != g.n {
		g.upload(arr)
	}
	n := len(arr)
	var res C.nb_step_result
	rc := C.nb_step(g.h, C.double(timeScaling), C.double(R), C.NB_STEP_DEFAULT, &res)
	if rc == C.NB_ERR_PAIR_OVERFLOW {
		// the step was NOT applied; the reference's event list has no capacity, so grow and run it again
		log.Printf("[INFO] event capacity %d exceeded: growing\n", g.pairCap)
		g.hostStale = true
		g.SyncToHost(bc)
		bigger := 4 * g.pairCap
		if m := 16*int64(n) + 65536; m > bigger {
			bigger = m
		}
		if err := g.recreate(g.cap, bigger); err != nil {
			log.Fatalf("[ERROR] %v", err)
		}
		g.upload(arr)
		rc = C.nb_step(g.h, C.double(timeScaling), C.double(R), C.NB_STEP_DEFAULT, &res)
	}
	if rc != C.NB_OK {
		log.Printf("[ERROR] nb_step: %s", g.lastError())
		g.Failed++
		return false
	}
	g.hostStale = true
	g.grow(n)
	// The device resolved the whole event queue in the reference's order (ProcessMods): elastic collisions,
	// ResolveSubsume and the `fragmenting` flag.  The records keep the Go bodies in step.
	if res.n_host_events > 0 {
		ev := make([]C.nb_event, int(res.n_host_events))
		var m C.int64_t
		C.nb_get_host_events(g.h, &ev[0], C.int64_t(len(ev)), &m)
		nFragInit, anySubsume := 0, false
		for _, e := range ev[:int(m)] {
			if e.kind == C.NB_EV_FRAG_INIT {
				nFragInit++
			}
			anySubsume = anySubsume || (e.kind == C.NB_EV_SUBSUME && e.applied != 0)
		}
		if anySubsume {
			if rc := C.nb_download_state(g.h, nil, nil, nil, nil, nil, nil, dptr(g.mass), nil, nil, nil,
				bptr(g.flags)); rc != C.NB_OK {
				log.Printf("[ERROR] nb_download_state: %s", g.lastError())
				return false
			}
		}
		// initiateFragmentation records where the body was while the queue was processed (fragcalc.go:77): the
		// position the cycle started from, not the one Update produced.  Many records: one bulk read.
		bulkPos := nFragInit > 16
		if bulkPos {
			if rc := C.nb_get_cycle_top_positions(g.h, 0, C.int64_t(n), dptr(g.x), dptr(g.y), dptr(g.z)); rc != C.NB_OK {
				log.Printf("[ERROR] nb_get_cycle_top_positions: %s", g.lastError())
				return false
			}
		}
		for _, e := range ev[:int(m)] {
			a, b := int(e.a), int(e.b)
			if a < 0 || b < 0 || a >= n || b >= n {
				continue
			}
			switch {
			case e.kind == C.NB_EV_SUBSUME && e.applied != 0:
				if arr[b].Exists && g.flags[b]&C.NB_F_EXISTS == 0 { // body.go:243
					log.Printf("[INFO] Body ID %v (mass %v) subsumed ID %v (mass %v)\n", arr[a].Id, arr[a].Mass,
						arr[b].Id, arr[b].Mass)
				}
				arr[a].Mass, arr[b].Mass = g.mass[a], g.mass[b]
				if g.flags[a]&C.NB_F_EXISTS == 0 {
					arr[a].Exists = false
				}
				if g.flags[b]&C.NB_F_EXISTS == 0 {
					arr[b].Exists = false
				}
			case e.kind == C.NB_EV_FRAG_INIT:
				// in the reference's handling order (the library sorts them): a body named twice keeps the later call
				var px, py, pz C.double
				if bulkPos {
					px, py, pz = C.double(g.x[a]), C.double(g.y[a]), C.double(g.z[a])
				} else if rc := C.nb_get_cycle_top_positions(g.h, C.int64_t(a), 1, &px, &py, &pz); rc != C.NB_OK {
					continue
				}
				arr[a].InitiateFragmentationAt(float64(e.f1), float64(e.dist), float64(px), float64(py), float64(pz))
			}
		}
	}
	// Renderables from the float32 snapshot — computation-runner.go:317-320
	xyz := unsafe.Slice((*float32)(unsafe.Pointer(g.pinXyz)), 3*n)
	ex := unsafe.Slice((*uint8)(unsafe.Pointer(g.pinExists)), n)
	for i, b := range arr {
		if b.Exists && ex[i] == 0 {
			log.Printf("[ERROR] NaN values. id=%d (removing from sim)", b.Id) // body.go:135
			b.Exists = false
		}
		r := body.NewRenderable(b) // Id, Radius, IsSun, Intensity, BodyColor (renderable.go:22-40)
		if b.Exists {
			r.X, r.Y, r.Z = xyz[3*i], xyz[3*i+1], xyz[3*i+2]
		}
		rq.Add(r)
	}
	return true
}

// Reserve: call before Cycle appends.  When `count` bodies will not fit, the Go bodies are refreshed while
// host and device indices still correspond; the next step re-creates the handle with room to spare.
func (g *GpuStepper) Reserve(bc *body.BodyCollection, count int) {
	if count <= g.cap {
		return
	}
	g.SyncToHost(bc)
	g.dirty = true
}

// AfterCycle keeps the device array in step with BodyCollection.Cycle (body_collection.go:253-296) without a
// re-upload: deaths are compacted on the device exactly like Cycle compacted the Go array (stable), adds are
// appended with r = R.
func (g *GpuStepper) AfterCycle(bc *body.BodyCollection, R float64) {
	if g.dirty {
		return
	}
	arr := bc.GetArray()
	var nDev C.int64_t
	if rc := C.nb_compact(g.h, &nDev, nil, 0); rc != C.NB_OK {
		g.dirty = true
		return
	}
	adds := len(arr) - int(nDev)
	if adds < 0 || len(arr) > g.cap {
		g.dirty = true
		return
	}
	if adds > 0 {
		g.grow(adds)
		for k := 0; k < adds; k++ {
			g.stage(k, arr[int(nDev)+k])
		}
		if rc := C.nb_append(g.h, C.int64_t(adds), C.double(R), dptr(g.x), dptr(g.y), dptr(g.z), dptr(g.vx),
			dptr(g.vy), dptr(g.vz), dptr(g.mass), dptr(g.radius), dptr(g.ff), dptr(g.fs), bptr(g.beh),
			bptr(g.flags)); rc != C.NB_OK {
			g.dirty = true
			return
		}
	}
	g.n = len(arr)
}
