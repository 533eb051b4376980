"""Debug aid: compress a few chunks on the GPU, parse the blocks in Python, report the first inconsistency."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lzbench_b200 as L
from tests import oracle_lib as O

def parse(blk, n, plain):
    ip = 0; op = 0; seqs = []
    out = bytearray()
    while True:
        if ip >= len(blk): return seqs, out, "token oob at ip=%d op=%d" % (ip, op)
        t = blk[ip]; ip0 = ip; ip += 1; l = t >> 4
        if l == 15:
            while True:
                if ip >= len(blk): return seqs, out, "litlen oob"
                b = blk[ip]; ip += 1; l += b
                if b != 255: break
        if ip + l > len(blk): return seqs, out, "literals oob: ip=%d l=%d clen=%d op=%d tokpos=%d" % (ip, l, len(blk), op, ip0)
        out += bytes(blk[ip:ip + l]); ip += l; op += l
        if ip == len(blk):
            seqs.append((ip0, op - l, l, 0, 0)); return seqs, out, None
        if ip + 2 > len(blk): return seqs, out, "offset oob"
        off = blk[ip] | (blk[ip + 1] << 8); ip += 2; m = t & 15
        if m == 15:
            while True:
                if ip >= len(blk): return seqs, out, "mlen oob"
                b = blk[ip]; ip += 1; m += b
                if b != 255: break
        m += 4
        if off == 0 or off > op: return seqs, out, "bad offset %d at op=%d tokpos=%d" % (off, op, ip0)
        for i in range(m): out.append(out[op - off + i])
        seqs.append((ip0, op - l, l, off, m)); op += m

kinds = sys.argv[1].split(",") if len(sys.argv) > 1 else ["random", "mixed"]
nchunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = L.Lz4B200(0)
for kind in kinds:
    data = L.corpus(kind, nchunks * 65536, seed=1)
    cs = L.chunk_list(len(data), 65536)
    packed, compr, total = ctx.compress_batch(data, cs)
    off = 0
    for i, z in enumerate(compr.tolist()):
        z = int(z); blk = packed[off:off + z]; plain = data[i * 65536:(i + 1) * 65536]
        v = O.oracle_validate(blk, 65536)
        ref = len(O.oracle_compress(plain))
        seqs, out, err = parse(blk.tolist(), 65536, plain)
        outa = np.frombuffer(bytes(out), dtype=np.uint8)
        m = min(len(outa), len(plain))
        neq = np.nonzero(outa[:m] != plain[:m])[0]
        first_bad = int(neq[0]) if len(neq) else -1
        print(f"{kind} chunk {i}: clen={z} ref={ref} validate={v} nseq={len(seqs)} decoded={len(out)} parse_err={err} first_mismatch={first_bad}")
        if v != 0 or err or first_bad >= 0 or len(out) != 65536:
            # show sequences around the problem
            pos = first_bad if first_bad >= 0 else len(out)
            near = [s for s in seqs if s[1] <= pos + 64 and s[1] + s[2] + s[4] >= pos - 200]
            print("   sequences near output pos", pos, "(tokpos, outpos, lit, off, mlen):")
            for s in near[-14:]: print("     ", s)
            print("   last 5 sequences:", seqs[-5:])
            break
        off += z
