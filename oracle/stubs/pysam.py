"""Empty stand-in for pysam: the reference imports it at package import time (mchap/io/loci.py:4) but
the hot path timed by bench.py's reference arm (DenovoMCMC.fit, CallingMCMC.fit, calling.exact) never
touches it.  Test infrastructure only."""
