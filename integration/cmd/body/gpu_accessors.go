package body

// gpu_accessors.go — the exported accessors the GPU stepper (cmd/runner/gpustepper.go) and the parity dump
// (cmd/sim/parity_dump_test.go) need.  No behaviour change: each one reads or writes an unexported field of
// Body / BodyCollection that body.go:34-52 and body_collection.go:14-39 already define.
//
// Drop this file into cmd/body of aceeric/nbodygo.

// IsFragmenting reports Body.fragmenting (body.go:45): such a body is skipped as i and as j by Compute
// (body.go:152-155,162-165) and keeps applying the force of its last Compute in Update.
func (b *Body) IsFragmenting() bool { return b.fragmenting }

// Restitution returns Body.r (body.go:44), the coefficient doElastic uses (collisioncalc.go:26-35).
func (b *Body) Restitution() float64 { return b.r }

// SetRestitution sets Body.r; Update does the same with the runner's R (body.go:129).
func (b *Body) SetRestitution(r float64) { b.r = r }

// Forces returns Body.fx, fy, fz (body.go:49): the force accumulated by the last Compute.
func (b *Body) Forces() (fx, fy, fz float64) { return b.fx, b.fy, b.fz }

// SetForces overwrites Body.fx, fy, fz — the GPU stepper mirrors the device's forces into the bodies that
// are fragmenting before it re-uploads a collection (nb_set_forces).
func (b *Body) SetForces(fx, fy, fz float64) { b.fx, b.fy, b.fz = fx, fy, fz }

// HasCollided reports Body.collided (body.go:51), set by doElastic and cleared by Update.
func (b *Body) HasCollided() bool { return b.collided }

// ClearCollided clears Body.collided like Update does (body.go:128).
func (b *Body) ClearCollided() { b.collided = false }

// DoFragment runs doFragment (fragcalc.go:54-61) with the factors shouldFragment produced.  On the GPU path
// the device evaluates shouldFragment in event order and reports (thisFactor, otherFactor) in an
// NB_EV_FRAGMENT record; the fragInfo bookkeeping stays here.
func (b *Body) DoFragment(otherBody *Body, thisFactor, otherFactor float64) {
	b.doFragment(otherBody, thisFactor, otherFactor)
}

// InitiateFragmentationAt runs initiateFragmentation (fragcalc.go:66-83) on the Mass and position the body had
// when its event was handled.  On the GPU path the device resolves the whole event queue; an NB_EV_FRAG_INIT
// record carries the mass of that moment (after the subsumes handled earlier in the queue), and
// nb_get_cycle_top_positions the position before Update.  The reference's own routine does the bookkeeping.
func (b *Body) InitiateFragmentationAt(fragFactor, massThen, x, y, z float64) {
	m, px, py, pz := b.Mass, b.X, b.Y, b.Z
	b.Mass, b.X, b.Y, b.Z = massThen, x, y, z
	b.initiateFragmentation(fragFactor)
	b.Mass, b.X, b.Y, b.Z = m, px, py, pz
}

// Fragment runs fragment (fragcalc.go:90-117): spawns the next batch of fragments as add events.
func (b *Body) Fragment(bc *BodyCollection) { b.fragment(bc) }

// EventRecord is one entry of the collection's deferred-event list (event.go:29-33) with the bodies named by
// their array index at the time of the call (-1: not in the array).
type EventRecord struct {
	Kind int // 0 collision, 1 subsume, 2 add (event.go:20-24)
	A, B int // indices of b1, b2 (add events: A = -1, B = -1)
}

// EventBacklog is the number of events still travelling through the channel between Enqueue and the
// handleEvents goroutine (body_collection.go:82-104).
func (bc *BodyCollection) EventBacklog() int { return len(bc.evCh) }

// PendingEvents returns the deferred-event list in the order ProcessMods will handle it (Front→Next,
// body_collection.go:212-233), without consuming it.
func (bc *BodyCollection) PendingEvents() []EventRecord {
	bc.lock.Lock()
	defer bc.lock.Unlock()
	index := make(map[*Body]int, len(bc.arr))
	for i, b := range bc.arr {
		index[b] = i
	}
	lookup := func(b *Body) int {
		if i, ok := index[b]; ok {
			return i
		}
		return -1
	}
	out := make([]EventRecord, 0, bc.events.Len())
	for e := bc.events.Front(); e != nil; e = e.Next() {
		ev := e.Value.(event)
		rec := EventRecord{Kind: int(ev.evType), A: -1, B: -1}
		if ev.evType != addEvent {
			rec.A, rec.B = lookup(ev.twoBodies.b1), lookup(ev.twoBodies.b2)
		}
		out = append(out, rec)
	}
	return out
}

// PendingAdds is countAdds (body_collection.go:236-244) under the lock.
func (bc *BodyCollection) PendingAdds() int {
	bc.lock.Lock()
	defer bc.lock.Unlock()
	return bc.countAdds()
}

// HasPendingRequests reports whether a get-body or mod-body request is waiting for the cycle top
// (body_collection.go:106-189): the GPU runner refreshes the Go bodies from the device before serving it.
func (bc *BodyCollection) HasPendingRequests() bool {
	return len(bc.getBodyCh) > 0 || len(bc.modBodyCh) > 0
}
