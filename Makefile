# Builds the product library (CUDA kernels + C ABI + lzbench wrappers + corpus generator)
# in-tree for sm_100a only.  `make oracle` builds the test-only checker under oracle/.
NVCC     ?= nvcc
CXX      ?= g++
CC       ?= gcc
ARCH      = -gencode arch=compute_100a,code=sm_100a
NVFLAGS   = $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
LIBDIR    = lzbench_b200/_lib
SRC       = lzbench_b200/csrc
HDRS      = $(wildcard $(SRC)/*.cuh) include/lz4b200.h

all: $(LIBDIR)/liblz4b200.so

$(LIBDIR)/lz4b200.o: $(SRC)/lz4b200.cu $(HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(LIBDIR)/ptxas.log || (cat $(LIBDIR)/ptxas.log; false)

$(LIBDIR)/lzbench_plugin.o: $(SRC)/lzbench_plugin.cpp include/lz4b200.h
	@mkdir -p $(LIBDIR)
	$(CXX) -std=c++11 -O2 -fPIC -Wall -Wextra -c -o $@ $<

$(LIBDIR)/corpus.o: $(SRC)/corpus.c
	@mkdir -p $(LIBDIR)
	$(CC) -O2 -fPIC -Wall -Wextra -c -o $@ $<

$(LIBDIR)/liblz4b200.so: $(LIBDIR)/lz4b200.o $(LIBDIR)/lzbench_plugin.o $(LIBDIR)/corpus.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -lpthread

oracle:
	$(MAKE) -C oracle

sass: $(LIBDIR)/liblz4b200.so
	@mkdir -p profiles
	cuobjdump -sass $(LIBDIR)/liblz4b200.so > profiles/sass_all.txt

clean:
	rm -rf $(LIBDIR)
	$(MAKE) -C oracle clean

.PHONY: all oracle sass clean
