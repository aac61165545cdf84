"""CPU reference arm: the UNMODIFIED reference package run through its own public API.

Test / measurement infrastructure only (bench.py's `cpu_baseline` leg and `--impl reference`); the
product never imports this module.

The reference is pure Python + numba, so there is nothing to compile with gcc: `build_ref()` (called
by `__graft_entry__.build()` in the build container, where /root/reference exists) pip-installs it
unchanged into `oracle/_ref/` (git-ignored, travels to the GPU box with the snapshot).  Its only
missing dependency on this image is pysam, imported at package import time by mchap/io/loci.py:4 and
never used by the functions timed here; `oracle/stubs/pysam.py` is an empty stand-in.

Parallelism is the reference's own `--cores` scheme (mchap/application/baseclass.py:360-388): a
`multiprocessing.Pool(n_cores)` over `np.array_split(items, n_cores)`; every worker calls
`DenovoMCMC.fit` / `CallingMCMC.fit` / `exact.posterior_mode` exactly like the CLIs do
(application/assemble.py:123-143, call.py:134-148, call_exact.py:126-172).  numba's JIT is warmed
(and cached under oracle/_ref/numba_cache) before the timed region.
"""
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")
REFERENCE_SRC = "/root/reference"


def build_ref(force=False):
    """pip-install the reference into oracle/_ref (only where /root/reference exists)."""
    if os.path.isdir(os.path.join(REF_DIR, "mchap")) and not force:
        return REF_DIR
    if not os.path.isdir(REFERENCE_SRC):
        return None
    import shutil
    import tempfile

    tmp = tempfile.mkdtemp()
    src = os.path.join(tmp, "reference")
    shutil.copytree(REFERENCE_SRC, src)   # the build writes egg-info next to the sources
    subprocess.check_call([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                           "--find-links", "/opt/wheelhouse", "--target", REF_DIR, "--upgrade", src],
                          stdout=subprocess.DEVNULL)
    shutil.rmtree(tmp, ignore_errors=True)
    return REF_DIR


def available():
    if not os.path.isdir(os.path.join(REF_DIR, "mchap")):
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


def _init():
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(REF_DIR, "numba_cache"))
    for p in (STUBS, REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings

    warnings.simplefilter("ignore")


# ------------------------------------------------------------------------------- workers
def _assemble_chunk(job):
    _init()
    from mchap.assemble import DenovoMCMC

    items, kw = job
    model = DenovoMCMC(**kw)
    n = 0
    for reads, counts in items:
        model.fit(reads, read_counts=counts)
        n += 1
    return n


def _call_mcmc_chunk(job):
    _init()
    from mchap.calling.classes import CallingMCMC

    items, kw = job
    n = 0
    for reads, counts, haps in items:
        CallingMCMC(haplotypes=haps, **kw).fit(reads, read_counts=counts)
        n += 1
    return n


def _call_exact_chunk(job):
    _init()
    from mchap.calling.exact import posterior_mode

    items, kw = job
    n = 0
    for reads, counts, haps in items:
        posterior_mode(reads, kw["ploidy"], haps, read_counts=counts, prior=kw["prior"], return_support_prob=True,
                       return_posterior_frequencies=True, return_posterior_occurrence=True)
        n += 1
    return n


_WORKERS = {"assemble": _assemble_chunk, "call_mcmc": _call_mcmc_chunk, "call_exact": _call_exact_chunk}


class Runner(object):
    """A warmed multiprocessing.Pool(cores) for one kind of work: `rate(items)` times one pass."""

    def __init__(self, kind, kw, cores, warm_item):
        import multiprocessing as mp

        self.worker, self.kw = _WORKERS[kind], dict(kw)
        warm_kw = dict(kw)
        if "steps" in warm_kw:
            warm_kw["steps"] = 5
        self.worker(([warm_item], warm_kw))   # compiles (or loads) in the parent and fills the on-disk cache
        self.cores = max(1, int(cores))
        ctx = mp.get_context("spawn")         # the parent may hold a CUDA context: never fork it
        self.pool = ctx.Pool(self.cores)
        self.pool.map(self.worker, [([warm_item], warm_kw)] * self.cores)   # every worker loads the cached code

    def rate(self, items):
        """(items per second, seconds) of one pass over `items`, split like np.array_split(items, cores)."""
        parts = [[items[i] for i in idx] for idx in np.array_split(np.arange(len(items)), self.cores) if len(idx)]
        t0 = time.perf_counter()
        done = sum(self.pool.map(self.worker, [(p, self.kw) for p in parts], chunksize=1))
        dt = time.perf_counter() - t0
        return done / dt, dt

    def close(self):
        self.pool.close()
        self.pool.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def pool_rate(kind, items, kw, cores):
    """Items per second of `kind` over `cores` processes (JIT warmed before the timed region).
    items: list of tuples as the worker of `kind` takes them; kw: model keyword arguments."""
    cores = max(1, min(int(cores), len(items)))
    with Runner(kind, kw, cores, items[0]) as r:
        rate, dt = r.rate(items)
    return rate, dt, cores
