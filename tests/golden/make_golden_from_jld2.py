"""Extract golden vectors for the encoding path from the reference's own serialized output.

Source: /root/reference/test/Data/ecg200/mps_saves/test_dataset.jld2 -- a JLD2 (HDF5 subset)
dump of a `TrainedMPS` (ECG200, Legendre_No_Norm, d=5).  No HDF5 reader exists in this image, so
the two arrays we need are located by content:

* `train_data.original_data` (Matrix{Float64}, 100 x 96, column-major, rows class-sorted as
  `encode_dataset` leaves them, encodings.jl:43-45): the first contiguous run of 9600 finite
  doubles in the file;
* the 9600 `PState.pstate[j]` vectors (Vector{Float64} of length d=5, stored sample-major then
  site): every one starts with the normalised P_0 = 0.7071067811865476.

Run here (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden_from_jld2.py
writes tests/golden/ecg200_legendre.npz  (X_orig (100,96), phi_ref (100,96,5)).
"""
import os
import re
import struct

import numpy as np

SRC = "/root/reference/test/Data/ecg200/mps_saves/test_dataset.jld2"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ecg200_legendre.npz")


def main():
    b = open(SRC, "rb").read()
    arr = np.frombuffer(b[: len(b) // 8 * 8], dtype="<f8")
    ok = np.isfinite(arr) & (np.abs(arr) < 1e3) & (np.abs(arr) > 1e-8)
    edges = np.diff(np.concatenate([[0], ok.astype(np.int8), [0]]))
    starts, ends = np.where(edges == 1)[0], np.where(edges == -1)[0]
    runs = [(int(s), int(e - s)) for s, e in zip(starts, ends) if e - s == 9600]
    assert len(runs) >= 1, runs
    X_orig = arr[runs[0][0]: runs[0][0] + 9600].reshape(96, 100).T.copy()      # (N=100, T=96)
    pat = struct.pack("<d", 0.7071067811865476)
    idx = [m.start() for m in re.finditer(re.escape(pat), b)]
    assert len(idx) == 9600, len(idx)
    phi = np.stack([np.frombuffer(b[i: i + 40], dtype="<f8") for i in idx]).reshape(100, 96, 5)
    np.savez_compressed(OUT, X_orig=X_orig, phi_ref=phi)
    print("wrote", OUT, X_orig.shape, phi.shape)


if __name__ == "__main__":
    main()
