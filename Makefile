# Convenience targets; everything they do is also reachable through __graft_entry__.py / pytest / bench.py.
PY ?= python

.PHONY: build test test-gpu bench clean

build:            ## libipn_b200.so (nvcc, sm_100a) + the C oracle (gcc)
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU suite: oracles, ABI, host logic, gloo world-size-2
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build   ## parity through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

bench: build      ## one JSON line (needs a B200)
	$(PY) bench.py

clean:
	rm -rf pyipn_b200/csrc/build pyipn_b200/libipn_b200.so oracle/_build tests/emu/_build
