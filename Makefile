# Top-level entry points.  Native code is built IN-TREE (k8s_cc_manager_b200/libccm.so + ccm-scrub,
# oracle/_build/libscrub_oracle.so); image targets live in deployments/container/Makefile
# (layout of the reference: Makefile -> deployments/container/Makefile, versions.mk).
include $(CURDIR)/versions.mk

PYTHON ?= python3

.PHONY: all native oracle test test-gpu smoke bench clean image

all: native oracle

native:            ## libccm.so (sm_100a) + ccm-scrub, next to the Python package
	$(PYTHON) -m k8s_cc_manager_b200.build

oracle:            ## CPU restatement of the scrub contract (test infrastructure only)
	$(MAKE) -s -C oracle

test: all          ## everything that runs without a GPU
	$(PYTHON) -m pytest tests/ -x -q -m "not gpu"

test-gpu: all      ## parity tests proper, on a B200
	$(PYTHON) -m pytest tests/ -x -q -m gpu

smoke: all
	$(PYTHON) -c "import __graft_entry__ as g; g.build(); g.smoke()"

bench: all
	$(PYTHON) bench.py

image:             ## distroless image (needs docker + network; not available in the dev sandbox)
	$(MAKE) -f deployments/container/Makefile build-distroless

clean:
	rm -f k8s_cc_manager_b200/libccm.so k8s_cc_manager_b200/ccm-scrub
	$(MAKE) -s -C oracle clean
