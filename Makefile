# Build everything in-tree.  The .so files are git-ignored but travel to the GPU box.
#   make            -> product (CUDA core + C ABI), host mirror, oracle
#   make product | host | oracle
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := /usr/bin/g++
PKG       := raytracer-2025_b200
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -fmad=false / -ffp-contract=off: the reference is strict IEEE binary64 without contraction
EXTRA     ?=
LIBNAME   ?= librt2025.so
NVCCFLAGS := $(EXTRA) -ccbin /usr/bin/g++ $(ARCH) -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fopenmp,-O3 -Iinclude -I$(PKG)/csrc
CXXFLAGS  := -O2 -std=c++17 -fPIC -ffp-contract=off -fopenmp -Wall -Wno-unknown-pragmas -Iinclude

PRODUCT_SRC := $(wildcard $(PKG)/csrc/*.cu) $(wildcard $(PKG)/csrc/*.cpp)
PRODUCT_HDR := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.h) include/rt2025.h include/rt2025_rng.h

all: product host oracle examples

product: $(PKG)/$(LIBNAME)
host: $(PKG)/librt2025_host.so
oracle: oracle/liboracle.so

$(PKG)/$(LIBNAME): $(PRODUCT_SRC) $(PRODUCT_HDR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(filter %.cu %.cpp,$(PRODUCT_SRC)) -Xlinker -lgomp

$(PKG)/librt2025_host.so: $(PKG)/host/host_capi.cpp $(PKG)/host/rt2025.hpp $(PKG)/host/scenes.hpp include/rt2025.h
	$(CXX) $(CXXFLAGS) -shared -o $@ $(PKG)/host/host_capi.cpp

oracle/liboracle.so: oracle/oracle.cpp include/rt2025.h include/rt2025_rng.h
	$(CXX) $(CXXFLAGS) -O3 -shared -o $@ oracle/oracle.cpp

examples: examples/final_scene

examples/final_scene: examples/final_scene.cpp $(PKG)/host/rt2025.hpp $(PKG)/host/scenes.hpp $(PKG)/librt2025.so
	$(CXX) $(CXXFLAGS) -o $@ examples/final_scene.cpp -L$(PKG) -lrt2025 -Wl,-rpath,'$$ORIGIN/../$(PKG)'

# host-only checks and tools of the scene compiler (no CUDA runtime needed)
HOSTTOOL_SRC := $(PKG)/csrc/compile.cpp $(PKG)/csrc/bvh_build.cpp
build/check_tie_order: scripts/check_tie_order.cpp $(HOSTTOOL_SRC) $(PRODUCT_HDR)
	mkdir -p build && $(CXX) $(CXXFLAGS) -I$(PKG)/csrc -I/usr/local/cuda/include -o $@ scripts/check_tie_order.cpp $(HOSTTOOL_SRC)
build/time_compile: scripts/time_compile.cpp $(HOSTTOOL_SRC) $(PKG)/host/host_capi.cpp $(PRODUCT_HDR)
	mkdir -p build && $(CXX) $(CXXFLAGS) -O3 -I$(PKG)/csrc -I$(PKG)/host -I/usr/local/cuda/include -o $@ scripts/time_compile.cpp $(HOSTTOOL_SRC) $(PKG)/host/host_capi.cpp
check_tie_order: build/check_tie_order
	build/check_tie_order
build/check_walk_entries: scripts/check_walk_entries.cpp $(HOSTTOOL_SRC) $(PKG)/host/host_capi.cpp $(PRODUCT_HDR)
	mkdir -p build && $(CXX) $(CXXFLAGS) -I$(PKG)/csrc -I$(PKG)/host -I/usr/local/cuda/include -o $@ scripts/check_walk_entries.cpp $(HOSTTOOL_SRC) $(PKG)/host/host_capi.cpp
check_walk_entries: build/check_walk_entries
	build/check_walk_entries
build/fuzz_compile: scripts/fuzz_compile.cpp $(HOSTTOOL_SRC) $(PKG)/host/host_capi.cpp $(PRODUCT_HDR)
	mkdir -p build && $(CXX) -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -std=c++17 -ffp-contract=off -Iinclude -I$(PKG)/csrc -I$(PKG)/host -I/usr/local/cuda/include -o $@ scripts/fuzz_compile.cpp $(HOSTTOOL_SRC) $(PKG)/host/host_capi.cpp
fuzz_compile: build/fuzz_compile
	ASAN_OPTIONS=detect_leaks=0 build/fuzz_compile 4000 1

clean:
	rm -f $(PKG)/librt2025.so $(PKG)/librt2025_host.so oracle/liboracle.so

.PHONY: all product host oracle examples clean check_tie_order check_walk_entries fuzz_compile
