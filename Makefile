# Builds the C-ABI shared library (portfft_b200/lib/libpfft_b200.so) for sm_100a, the C oracle and, when the
# reference checkout is present, the reference-shim library under oracle/_ref/.
NVCC      ?= nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unknown-pragmas -diag-suppress 20013
CSRC      := portfft_b200/csrc
BUILD     := build
LIB       := portfft_b200/lib/libpfft_b200.so
CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CPP_SRCS  := $(wildcard $(CSRC)/*.cpp)
OBJS      := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(CU_SRCS)) $(patsubst $(CSRC)/%.cpp,$(BUILD)/%.o,$(CPP_SRCS))
HDRS      := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/pfft.h

all: $(LIB) oracle build/api_smoke

$(BUILD)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(BUILD)/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(BUILD)
	$(CXX) -std=c++17 -O2 -fPIC -Wall -Wno-unknown-pragmas -I/usr/local/cuda/include -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p portfft_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(BUILD) $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean

# C++ smoke test of the header-only API mirror (include/portfft/portfft.hpp); run on the GPU by tests/test_cpp_api.py
build/api_smoke: tests/cpp/api_smoke.cpp $(LIB) $(wildcard include/portfft/*.hpp)
	@mkdir -p build
	g++ -std=c++17 -O1 -Wall -Iinclude -I/usr/local/cuda/include $< -o $@ -Lportfft_b200/lib -lpfft_b200 \
	  -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/../portfft_b200/lib'
