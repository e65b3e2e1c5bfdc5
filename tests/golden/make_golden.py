"""Regenerates tests/golden/*.npz from the reference's own headers (oracle/_ref).

Run in the authoring container only (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The fixtures pin oracle/oracle_ife.cpp (and through it the CUDA path) to the outputs of
the unmodified reference code for inputs the reference's own tests do not cover: the
float instantiation of the solver/functor, degenerate and diagonal matrices, NaN and
on-edge histogram inserts, duplicate-heavy edge determination.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import oracle as O  # noqa: E402
import synth  # noqa: E402

O.build(force=True)
R = O.Ref()
assert R.math_overload_is_double()

A = synth.special_matrices(4000, seed=11)
np.savez_compressed(os.path.join(HERE, "solver_ref.npz"), A6=A,
                    features_f32=R.features_f32(A), eig_f32=R.eig_f32(A),
                    eig_f64=R.eig_f64(A.astype(np.float64)))

rng = np.random.default_rng(5)
edges = np.sort(rng.standard_normal(40)).astype(np.float32)
vals = np.concatenate([rng.standard_normal(5000).astype(np.float32) * 1.5, edges,
                       np.nextafter(edges, np.float32(np.inf)), np.nextafter(edges, np.float32(-np.inf)),
                       np.array([np.nan, np.inf, -np.inf, 0.0, -0.0], np.float32)])
counts, freqs = R.hist_f32(edges, vals)
np.savez_compressed(os.path.join(HERE, "hist_ref.npz"), edges=edges, values=vals, counts=counts,
                    freqs=freqs)

cases = {}
s1 = np.sort(rng.standard_normal(1000))
s2 = np.sort(rng.integers(0, 12, 500).astype(np.float64))          # many duplicates
s3 = np.sort(np.concatenate([np.zeros(300), rng.standard_normal(200)]))
for name, s, nb in (("normal41", s1, 41), ("dups7", s2, 7), ("zeros10", s3, 10), ("dups41", s2, 41)):
    cases[name + "_samples"] = s
    cases[name + "_edges"] = R.determine_edges(s, nb)
    cases[name + "_nbins"] = np.array(nb)
np.savez_compressed(os.path.join(HERE, "edges_ref.npz"), **cases)
print("golden fixtures written to", HERE)
