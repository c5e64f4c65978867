"""Extract the reference's own TRAINED MPS (ECG200, Legendre_No_Norm, d=5) from its serialized fixture.

Source: /root/reference/test/Data/ecg200/mps_saves/test_dataset.jld2 (JLD2 = HDF5 subset; no HDF5 reader in this
image).  An `ITensor` is stored as a compound datum
    [RelOffset of the Dense storage Vector{Float64}] [Index]*N
with Index = id::UInt64, space::Int64, dir::Int32, tags::(4 x UInt256, length::Int64), plev::Int64 (164 bytes).
Tag strings ("Site", "n=3", "Link", "l=2", "f(x)") sit one character per UInt16 in the top byte, most significant
first.  The storage vector is an HDF5 v2 object header ("OHDR") whose dataspace (rank-1 length) and contiguous layout
message give the data address; all addresses are relative to the 512-byte JLD2 base.

Writes tests/golden/ecg200_trained_mps.npz:
    core_XX  : (chi_l, d, chi_r[, C]) Float64 in this repo's python core layout, XX = site (0-based)
    label_pos: site carrying the class index "f(x)"
    class_distribution (from the training set stored next to it; the class-sorted rows of ecg200_legendre.npz X_orig)
Run here (the reference tree does not exist on the GPU box):  python tests/golden/make_golden_mps_from_jld2.py
"""
import os
import re
import struct

import numpy as np

SRC = "/root/reference/test/Data/ecg200/mps_saves/test_dataset.jld2"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ecg200_trained_mps.npz")
BASE = 512
ISZ = 164


def tagstr(f):
    return "".join(chr(f[31 - 2 * k]) for k in range(16) if f[31 - 2 * k])


def parse_index(b, o):
    id_, space, dirv = struct.unpack_from("<QqI", b, o)
    length, plev = struct.unpack_from("<qq", b, o + 20 + 128)
    tags = [tagstr(b[o + 20 + 32 * k: o + 20 + 32 * (k + 1)]) for k in range(4)]
    return dict(id=id_, space=space, dir=dirv, tags=tags[:max(0, min(length, 4))], ntags=length, plev=plev)


def plausible(ix):
    return 1 <= ix["space"] <= 4096 and 1 <= ix["ntags"] <= 4 and ix["plev"] == 0 and all(ix["tags"])


def read_f64_vector(b, rel):
    o = rel + BASE
    assert b[o:o + 4] == b"OHDR", (o, b[o:o + 8])
    flags = b[o + 5]
    p = o + 6
    if flags & 0x20:
        p += 16
    if flags & 0x10:
        p += 4
    nb = 1 << (flags & 3)
    chunk = int.from_bytes(b[p:p + nb], "little")
    p += nb
    end = p + chunk
    n = addr = size = None
    while p < end:
        t = b[p]
        sz = struct.unpack_from("<H", b, p + 1)[0]
        p += 4 + (2 if flags & 0x04 else 0)
        body = b[p:p + sz]
        if t == 1:
            assert body[1] == 1                                   # rank-1 dataspace
            n = struct.unpack_from("<Q", body, 4)[0]
        elif t == 8:
            assert body[0] in (3, 4) and body[1] == 1, body[:2]   # contiguous layout
            addr, size = struct.unpack_from("<QQ", body, 2)
        p += sz
    assert size == 8 * n
    return np.frombuffer(b[addr + BASE: addr + BASE + size], dtype="<f8").copy()


def main():
    b = open(SRC, "rb").read()
    site_pat = bytes([ord("e"), 0, ord("t"), 0, ord("i"), 0, ord("S")])
    cores = {}
    label_pos = None
    for m in re.finditer(re.escape(site_pat), b):
        o = m.start() - 45                                         # start of the Index whose first tag is "Site"
        ix0 = parse_index(b, o)
        if not plausible(ix0) or ix0["tags"][0] != "Site":
            continue
        # the ITensor's indices are contiguous; the Site index may not be the first: walk back, then forward
        first = o
        while plausible(parse_index(b, first - ISZ)):
            first -= ISZ
        inds = []
        q = first
        while plausible(parse_index(b, q)) and len(inds) < 4:
            inds.append(parse_index(b, q))
            q += ISZ
        rel = struct.unpack_from("<Q", b, first - 8)[0]
        data = read_f64_vector(b, rel)
        dims = [ix["space"] for ix in inds]
        assert data.size == int(np.prod(dims)), (dims, data.size)
        A = data.reshape(dims, order="F")                         # Julia column-major, first index fastest
        site = int([t for t in ix0["tags"] if t.startswith("n=")][0][2:]) - 1
        role = {}
        for ax, ix in enumerate(inds):
            if "Site" in ix["tags"]:
                role["s"] = ax
            elif "f(x)" in ix["tags"]:
                role["c"] = ax
            else:
                l = int([t for t in ix["tags"] if t.startswith("l=")][0][2:])
                role["a" if l == site else "b"] = ax              # Link l=j joins sites j and j+1 (1-based)
        # python layout (chi_l, d, chi_r[, C]) with missing boundary links as size-1 axes
        A = np.moveaxis(A, [role[k] for k in "asbc" if k in role], range(len(role)))
        have = [k for k in "asbc" if k in role]
        shape = []
        it = iter(A.shape)
        for k in "asb":
            shape.append(next(it) if k in have else 1)
        if "c" in have:
            shape.append(next(it))
            label_pos = site
        cores[site] = np.ascontiguousarray(A.reshape(shape))
    T = len(cores)
    assert sorted(cores) == list(range(T)) and label_pos is not None, (sorted(cores), label_pos)
    for j in range(T - 1):
        assert cores[j].shape[2] == cores[j + 1].shape[0], (j, cores[j].shape, cores[j + 1].shape)
    out = {"core_%02d" % j: cores[j] for j in range(T)}
    out["label_pos"] = np.int64(label_pos)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "T =", T, "label at", label_pos, "chi_max =", max(c.shape[2] for c in cores.values()))


if __name__ == "__main__":
    main()
