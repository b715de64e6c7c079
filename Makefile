# Builds the C-ABI library (sm_100a only) and the CPU oracle.  `python -c "import __graft_entry__ as g; g.build()"`
# runs the same commands.
NVCC ?= /usr/local/cuda/bin/nvcc
NVCCFLAGS = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
            -Xcompiler -fPIC -Xptxas -v
LIB = bayes_qnet_b200/libbq_b200.so
ORACLE = oracle/libbq_oracle.so

all: $(LIB) $(ORACLE)

SRCS = bayes_qnet_b200/csrc/bq_lib.cu bayes_qnet_b200/csrc/bq_batch_nw.cu
$(LIB): $(SRCS) $(wildcard bayes_qnet_b200/csrc/*.cuh) bayes_qnet_b200/csrc/bq_host.h include/bq_capi.h
	$(NVCC) $(NVCCFLAGS) --threads 2 -shared -o $@ $(SRCS) -lcudart

$(ORACLE): oracle/bq_oracle.c
	gcc -O2 -ffp-contract=off -fPIC -shared -o $@ oracle/bq_oracle.c -lm

clean:
	rm -f $(LIB) $(ORACLE)
