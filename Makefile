# Builds the sm_100a CUDA library (the product) and the CPU oracle (test infrastructure).
#   make            -> zk-paillier_b200/libzkp_b200.so + oracle/liboracle.so
#   make lib        -> CUDA library only
#   make oracle     -> oracle only
#   make examples   -> examples/range_proof_ni (the reference's range-proof test against the C++ mirror)
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
CSRC      := zk-paillier_b200/csrc
BUILD     := build
CU        := $(wildcard $(CSRC)/*.cu)
OBJ       := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(CU))
LIB       := zk-paillier_b200/libzkp_b200.so
HOSTLIB   := zk-paillier_b200/libzkp_host.so
HOSTSRC   := zk-paillier_b200/host

all: lib oracle

# lab build: the measured-and-rejected kernel variants, the pipe probes and the ZKP_B200_* environment knobs
# (scripts/k1m_variants.py, scripts/imad_probes.py with ZKP_B200_LIB=zk-paillier_b200/libzkp_b200_lab.so); not shipped
# variants of the lab build:  make lab LABTAG=_nosub LABFLAGS=-DZKP_B200_LAB_NOSUB  (timing probe with wrong results, see mp_coop.cuh)
LABTAG    ?=
LABFLAGS  ?=
LABBUILD  := build/lab$(LABTAG)
LABOBJ    := $(patsubst $(CSRC)/%.cu,$(LABBUILD)/%.o,$(CU))
LABLIB    := zk-paillier_b200/libzkp_b200_lab$(LABTAG).so
lab: $(LABLIB)
$(LABBUILD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/zkp_b200.h
	@mkdir -p $(LABBUILD)
	$(NVCC) $(NVFLAGS) -DZKP_B200_LAB $(LABFLAGS) -c $< -o $@ 2> $(LABBUILD)/$*.ptxas.log || (cat $(LABBUILD)/$*.ptxas.log; false)
$(LABLIB): $(LABOBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(LABOBJ) -lcudart

lib: $(LIB) $(HOSTLIB)

# C++ host mirror of the reference's zkproofs::* interface (JSON C shim for the tests); links the CUDA library
$(HOSTLIB): $(HOSTSRC)/host_capi.cpp $(wildcard $(HOSTSRC)/*.hpp) include/zkp_b200.h $(LIB)
	g++ -O2 -std=c++17 -fPIC -shared -Wall -o $@ $(HOSTSRC)/host_capi.cpp -Lzk-paillier_b200 -lzkp_b200 -Wl,-rpath,'$$ORIGIN'

$(BUILD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/zkp_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; false)

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

examples: examples/range_proof_ni
examples/range_proof_ni: examples/range_proof_ni.cpp $(wildcard $(HOSTSRC)/*.hpp) include/zkp_b200.h $(LIB)
	g++ -O2 -std=c++17 -Wall -o $@ $< -Lzk-paillier_b200 -lzkp_b200 -Wl,-rpath,'$$ORIGIN/../zk-paillier_b200'

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(BUILD) $(LIB) $(LABLIB) $(HOSTLIB) examples/range_proof_ni
	$(MAKE) -C oracle clean

.PHONY: all lib lab oracle examples clean
