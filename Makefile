# Build libgr2m_b200.so without Python (what an R / C / Fortran caller needs).  sm_100a only.
NVCC ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC
CSRC := gr2msemidistr_b200/csrc
LIB := gr2msemidistr_b200/libgr2m_b200.so

lib: $(LIB)

$(LIB): $(CSRC)/gr2m.cu $(CSRC)/wfa.cu $(CSRC)/d8prep.cu $(CSRC)/router.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.inc) include/gr2m_b200.h
	$(NVCC) $(NVCCFLAGS) -o $@ $(CSRC)/gr2m.cu $(CSRC)/wfa.cu $(CSRC)/d8prep.cu $(CSRC)/router.cu

oracle:
	$(MAKE) -C oracle

demo: $(LIB) examples/c_abi_demo.c
	$(CC) -std=c99 -Wall -Wextra -pedantic -Iinclude -o examples/c_abi_demo examples/c_abi_demo.c -Lgr2msemidistr_b200 -lgr2m_b200 -Wl,-rpath,'$$ORIGIN/../gr2msemidistr_b200' -lm

clean:
	rm -f $(LIB) examples/c_abi_demo
	$(MAKE) -C oracle clean

.PHONY: lib oracle demo clean
