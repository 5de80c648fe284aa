# Top-level convenience targets (the Python entry point __graft_entry__.build() does the same).
NVCC    ?= /usr/local/cuda/bin/nvcc
CSRC    := nbodygo_b200/csrc
SO      := nbodygo_b200/libnbody_b200.so
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared

lib: $(SO)

$(SO): $(CSRC)/nb_force.cu $(CSRC)/nb_resolve.cu $(CSRC)/nb_integrate.cu $(CSRC)/nb_api.cu $(CSRC)/nb_internal.cuh include/nbody_b200.h
	$(NVCC) $(NVFLAGS) -o $@ $(CSRC)/nb_force.cu $(CSRC)/nb_resolve.cu $(CSRC)/nb_integrate.cu $(CSRC)/nb_api.cu -ldl

oracle:
	$(MAKE) -C oracle

test: lib oracle
	python -m pytest tests -x -q -m "not gpu"

clean:
	rm -f $(SO)
	$(MAKE) -C oracle clean

.PHONY: lib oracle test clean
