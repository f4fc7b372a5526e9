# Build / test / measure recipes (reference: Makefile install-* targets; there is nothing to pip-install here --
# the image already has torch 2.11 + CUDA 12.9, and Apex is replaced by ddpt_b200.amp).
PYTHON ?= python

.PHONY: build install-all train test test-gpu smoke bench bench-ref sass sanitize clean
NPROC ?= 1
PORT ?= 6666

build:            ## compile every CUDA/C++ source for sm_100a into <package>/_ext/_ddpt_C*.so
	$(PYTHON) -c "import __graft_entry__ as g; g.build()"

install-all: build  ## the reference's entry target (Makefile:7-8); nothing to download here

train:            ## the reference's run line (README.md:28): one process per GPU, run.py -n=DDP_warmup
	$(PYTHON) -m torch.distributed.run --nnodes=1 --nproc-per-node $(NPROC) --master-addr 127.0.0.1 --master-port $(PORT) run.py -n=DDP_warmup

test:             ## CPU suite (gloo world_size 2, oracle from the reference's shipped logs)
	$(PYTHON) -m pytest tests -q -m "not gpu"

test-gpu:         ## kernel numerics, fused Toy_Net, comm protocol, multi-GPU (needs a B200)
	$(PYTHON) -m pytest tests -q -m gpu

smoke:
	$(PYTHON) -c "import __graft_entry__ as g; g.smoke()"

bench:            ## one JSON line; N>1: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N
	$(PYTHON) bench.py --gpus 1 --steps 200 --warmup 10

bench-ref:        ## the unmodified reference sources on the same config
	bash baseline/install_ref.sh && $(PYTHON) bench.py --impl reference

sass:             ## SASS evidence (tcgen05 / TMA / multimem mnemonics per kernel) -> sass/
	$(PYTHON) benchmarks/dump_sass.py

sanitize:         ## memcheck + racecheck + synccheck over the single-GPU kernel tests
	for tool in memcheck racecheck synccheck; do \
	  compute-sanitizer --tool $$tool --error-exitcode 1 $(PYTHON) -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "not mxfp8" || exit 1; \
	done

clean:
	rm -rf pytorch-distributeddataparallel-training-tricks_b200/_ext/obj pytorch-distributeddataparallel-training-tricks_b200/_ext/*.so
