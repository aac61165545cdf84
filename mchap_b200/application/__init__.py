"""Batched application layer: the reference's ``mchap assemble`` / ``call`` / ``call-exact`` programs
re-plumbed so that blocks of loci go through the device in one call per stage
(SURVEY.md section 8(f) N3, N4)."""
from .programs import (LocusAssemblyError, SampleAssemblyError, assemble_program, call_exact_program,
                       call_program, call_posterior_haplotypes)
from .cli import build_program, main

__all__ = ["assemble_program", "call_program", "call_exact_program", "call_posterior_haplotypes",
           "LocusAssemblyError", "SampleAssemblyError", "build_program", "main"]
