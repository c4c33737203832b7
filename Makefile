# Builds the sm_100a device backend (C ABI in include/spn_b200.h), the C++ host
# layer and the BS executables.  nvcc cross-compiles without a GPU.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude $(SPN_TUNE)
CSRC      := streamparanumal_b200/csrc
LIBDIR    := streamparanumal_b200/lib

KERNEL_SRCS := $(CSRC)/runtime.cu $(CSRC)/stream_abi.cu $(CSRC)/ogs_abi.cu $(CSRC)/ogs_plan.cu $(CSRC)/ogs_halo.cu
KERNEL_HDRS := $(CSRC)/common.cuh $(CSRC)/stream_kernels.cuh $(CSRC)/ogs_kernels.cuh $(CSRC)/ogs_pipe.cuh $(CSRC)/ogs_tile.cuh include/spn_b200.h

all: $(LIBDIR)/libspn_b200.so $(LIBDIR)/libspn_b200_comm.so $(LIBDIR)/libspn_host.so bs

$(LIBDIR)/libspn_b200.so: $(KERNEL_SRCS) $(KERNEL_HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(KERNEL_SRCS) -cudart static

$(LIBDIR)/libspn_b200_comm.so: $(CSRC)/comm_nccl.cu include/spn_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/comm_nccl.cu -cudart static -lnccl

# C++ host layer (mirror of the reference's ogs / platform / comm API) + its C binding
HOST      := streamparanumal_b200/host
HOST_SRCS := $(HOST)/comm.cpp $(HOST)/platform.cpp $(HOST)/ogs_setup.cpp $(HOST)/ogs_ops.cpp $(HOST)/mesh.cpp $(HOST)/kernel.cpp $(HOST)/capi.cpp
HOST_HDRS := $(HOST)/comm.hpp $(HOST)/platform.hpp $(HOST)/ogs.hpp $(HOST)/kernel.hpp include/spn_b200.h include/spn_host.h
CXXFLAGS  := -O2 -g -std=c++17 -fPIC -fopenmp -Wall -Iinclude -I$(HOST)
$(LIBDIR)/libspn_host.so: $(HOST_SRCS) $(HOST_HDRS) $(LIBDIR)/libspn_b200.so $(LIBDIR)/libspn_b200_comm.so
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_SRCS) -L$(LIBDIR) -lspn_b200 -lspn_b200_comm -Wl,-rpath,'$$ORIGIN'

# BS0..BS8 executables (reference CLI and output lines) and the local rank launcher (mpirun stand-in)
BSDIR     := streamparanumal_b200/bs
BINDIR    := streamparanumal_b200/bin
BS_BINS   := $(foreach n,0 1 2 3 4 5 6 7 8,$(BINDIR)/BS$(n)) $(BINDIR)/spnrun
bs: $(BS_BINS)
BS_LINK   := -L$(LIBDIR) -lspn_host -lspn_b200 -lspn_b200_comm -Wl,-rpath,'$$ORIGIN/../lib'
$(BINDIR)/BS0 $(BINDIR)/BS1 $(BINDIR)/BS2 $(BINDIR)/BS3 $(BINDIR)/BS4 $(BINDIR)/BS5: $(BSDIR)/bs_stream.cpp $(BSDIR)/settings.hpp $(LIBDIR)/libspn_host.so
	@mkdir -p $(BINDIR)
	$(CXX) $(CXXFLAGS) -fPIE -I$(BSDIR) -DSPN_BS=$(subst $(BINDIR)/BS,,$@) -o $@ $(BSDIR)/bs_stream.cpp $(BS_LINK)
$(BINDIR)/BS6 $(BINDIR)/BS7 $(BINDIR)/BS8: $(BSDIR)/bs_ogs.cpp $(BSDIR)/settings.hpp $(LIBDIR)/libspn_host.so
	@mkdir -p $(BINDIR)
	$(CXX) $(CXXFLAGS) -fPIE -I$(BSDIR) -DSPN_BS=$(subst $(BINDIR)/BS,,$@) -o $@ $(BSDIR)/bs_ogs.cpp $(BS_LINK)
$(BINDIR)/spnrun: $(BSDIR)/spnrun.cpp
	@mkdir -p $(BINDIR)
	$(CXX) -O2 -std=c++17 -Wall -o $@ $(BSDIR)/spnrun.cpp

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o /dev/null $(CSRC)/stream_abi.cu 2>&1 | grep -E "Compiling|registers|spill"
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o /dev/null $(CSRC)/ogs_abi.cu 2>&1 | grep -E "Compiling|registers|spill"

clean:
	rm -rf $(LIBDIR)/*.so $(BINDIR) build

.PHONY: all clean ptxas-info bs

# ---- tuning tools (run on the GPU box) ----------------------------------------------
tune: build/tune_stream
build/tune_stream: tools/tune_stream.cu $(KERNEL_HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -o $@ tools/tune_stream.cu
.PHONY: tune

# ---- the reference's own sources compiled UNCHANGED against the shims (include/shim/{occa.hpp,mpi.h}) ----------------
# Proof of the drop-in boundary: libs/core, libs/primitives, libs/ogs, libs/mesh and every BS driver are compiled from
# where they lie in $(REF) — nothing is copied or edited — and linked with the product backend of the OCCA shim
# (shim/occa_spn.cpp -> C ABI -> sm_100a kernels), the single-rank MPI stand-in and the small LAPACK subset the hex
# reference element needs.  Built only where the reference tree exists; the binaries travel to the GPU box.
REF        ?= /root/reference
SHIM       := streamparanumal_b200/shim
REFDRV     := build/ref_drivers
REF_CXXFLAGS := -fopenmp -O3 -std=c++17 -fPIC -Iinclude/shim -Iinclude -I$(REF)/include \
               -DLIBP_DIR='"$(REF)"' -DOGS_DIR='"$(REF)/libs/ogs"' -DMESH_DIR='"$(REF)/libs/mesh"' -DSTREAM_DIR='"$(REF)"'
REF_LIB_SRCS := $(wildcard $(REF)/libs/core/*.cpp) $(wildcard $(REF)/libs/primitives/*.cpp) $(wildcard $(REF)/libs/ogs/*.cpp) \
               $(wildcard $(REF)/libs/mesh/*.cpp)
REF_LIB_OBJS := $(patsubst $(REF)/libs/%.cpp,$(REFDRV)/obj/%.o,$(REF_LIB_SRCS))
SHIM_OBJS    := $(REFDRV)/obj/shim/occa_spn.o $(REFDRV)/obj/shim/mpi_shim.o $(REFDRV)/obj/shim/minilapack.o

ref-drivers:
	@if [ -d $(REF)/BS ]; then $(MAKE) --no-print-directory $(foreach n,0 1 2 3 4 5 6 7 8,$(REFDRV)/BS$(n)) $(REFDRV)/ogs_check; \
	 else echo "ref-drivers: $(REF) absent, keeping prebuilt $(REFDRV) (if any)"; fi

$(REFDRV)/obj/%.o: $(REF)/libs/%.cpp include/shim/occa.hpp include/shim/mpi.h
	@mkdir -p $(dir $@)
	$(CXX) $(REF_CXXFLAGS) -c -o $@ $<
$(REFDRV)/obj/shim/%.o: $(SHIM)/%.cpp include/shim/occa.hpp include/shim/mpi.h include/spn_b200.h
	@mkdir -p $(dir $@)
	$(CXX) $(REF_CXXFLAGS) -c -o $@ $<
$(REFDRV)/libref_host_layer.a: $(REF_LIB_OBJS)
	ar rcs $@ $^

# the reference's ogs_t on the device against its own host path (tests/cpp/ref_ogs_device_check.cpp)
$(REFDRV)/ogs_check: tests/cpp/ref_ogs_device_check.cpp $(REFDRV)/libref_host_layer.a $(SHIM_OBJS) $(LIBDIR)/libspn_b200.so
	$(CXX) $(REF_CXXFLAGS) -fPIE -o $@ $< $(SHIM_OBJS) $(REFDRV)/libref_host_layer.a -L$(LIBDIR) -lspn_b200 -Wl,-rpath,'$$ORIGIN/../../$(LIBDIR)'

define REF_BS
$(REFDRV)/BS$(1): $(REF)/BS/BS$(1)/bs$(1)Main.cpp $(wildcard $(REF)/BS/BS$(1)/src/*.cpp) $(REFDRV)/libref_host_layer.a $(SHIM_OBJS) $(LIBDIR)/libspn_b200.so
	$(CXX) $(REF_CXXFLAGS) -fPIE -I$(REF)/BS/BS$(1) -o $$@ $(REF)/BS/BS$(1)/bs$(1)Main.cpp $(wildcard $(REF)/BS/BS$(1)/src/*.cpp) \
	  $(SHIM_OBJS) $(REFDRV)/libref_host_layer.a -L$(LIBDIR) -lspn_b200 -Wl,-rpath,'$$$$ORIGIN/../../$(LIBDIR)'
endef
$(foreach n,0 1 2 3 4 5 6 7 8,$(eval $(call REF_BS,$(n))))
.PHONY: ref-drivers
