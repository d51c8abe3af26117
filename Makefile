# Convenience targets; the names follow the reference's Makefile where they mean the same thing
# (reference Makefile:8-69: run / test / lib / clean).  Everything heavy lives in Python:
# eigen_value_b200/build.py is the nvcc recipe, oracle/Makefile the CPU oracle's.
PY ?= python

lib:            ## reference `make lib` (Makefile:66-69): build libsimilarity_transform.so for sm_100a
	$(PY) -m eigen_value_b200.build

oracle:         ## CPU oracle (test infrastructure)
	$(MAKE) -C oracle

ref:            ## the unmodified reference sources on the CPU SYCL shim (needs /root/reference)
	$(MAKE) -C oracle ref

build: lib oracle
	$(PY) -c "import __graft_entry__ as g; g.build()"

test:           ## CPU suite: oracle vs golden vectors / real reference, ABI, gloo loop
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:       ## parity through the C ABI on a B200
	$(PY) -m pytest tests -x -q -m gpu

test-emulated:  ## the -m gpu test files on the CPU, against the whole library built on the emulation harness
	ST_EMULATED_LIB=1 ST_EMU_DEVICES=4 $(PY) -m pytest tests -q -m gpu

racecheck:      ## the emulated round kernels under ThreadSanitizer (+ mutants that must be reported)
	$(PY) -m pytest tests/test_kernel_racecheck_emulated.py -q

run: bench      ## reference `make run` prints its benchmark table; here: one JSON line
bench:
	$(PY) bench.py

golden:         ## regenerate tests/golden/ from the reference (build container only)
	$(PY) tests/golden/make_golden.py
	$(PY) tests/golden/make_reference_golden.py

clean:
	rm -f eigen_value_b200/libsimilarity_transform.so tests/cpp/*.bin tools/stream_probe
	rm -rf tests/cuda_emu/_build tests/cuda_emu/*.so tests/cuda_emu/*.bin
	$(MAKE) -C oracle clean

.PHONY: lib oracle ref build test test-gpu test-emulated racecheck run bench golden clean
