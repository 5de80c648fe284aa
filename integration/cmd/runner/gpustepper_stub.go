//go:build !gpu

package runner

// gpustepper_stub.go — keeps cmd/runner building without libnbody_b200.so.  There is no CPU fallback behind
// --gpu: a binary built without the `gpu` tag refuses the flag instead of silently running the work pool.

import (
	"errors"
	"nbodygo/cmd/body"
)

type GpuStepper struct{ Failed uint }

func NewGpuStepper(device, capacity int) (*GpuStepper, error) {
	return nil, errors.New("this binary was built without -tags gpu (libnbody_b200.so)")
}

func (g *GpuStepper) Close()                                        {}
func (g *GpuStepper) MarkDirty()                                    {}
func (g *GpuStepper) SyncToHost(bc *body.BodyCollection)            {}
func (g *GpuStepper) Reserve(bc *body.BodyCollection, count int)    {}
func (g *GpuStepper) AfterCycle(bc *body.BodyCollection, R float64) {}
func (g *GpuStepper) Step(bc *body.BodyCollection, timeScaling, R float64, rq *ResultQueue) bool {
	return false
}
