"""Extracts the production verifier CRS embedded in the reference's browser verifier
(packages/backend-wasm/src/verifier/generated/sigma-verify.generated.ts) into tests/golden/verifier_crs_kat.json.

These are real production CRS points (ffjavascript encoding: Montgomery form with R = 2^384, little-endian limbs).  They
pin (a) the G1/G2 decoding and on-curve checks and (b) the pairing: e(sigma1.x, H) = e(G, sigma2.x) and the same for y.
Run in the build container only (reads /root/reference); the JSON it writes is the committed fixture."""
import json
import re
import sys

SRC = "/root/reference/packages/backend-wasm/src/verifier/generated/sigma-verify.generated.ts"
Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
RINV = pow(1 << 384, Q - 2, Q)


def fq(b):
    return int.from_bytes(bytes(b), "little") * RINV % Q


def main():
    text = open(SRC).read()
    arrays = {}
    stack = []
    for line in text.splitlines():
        m = re.match(r"\s*([A-Za-z0-9_]+): \{\s*$", line)
        if m:
            stack.append(m.group(1))
            continue
        if re.match(r"\s*\},?\s*$", line) and stack:
            stack.pop()
            continue
        m = re.match(r"\s*([A-Za-z0-9_]+): Uint8Array\.from\(\[([0-9,]*)\]\)", line)
        if m:
            arrays[".".join(stack + [m.group(1)])] = [int(x) for x in m.group(2).split(",")]
    out = {}
    for k, b in arrays.items():
        if len(b) == 96:
            out[k] = {"group": "G1", "x": hex(fq(b[:48])), "y": hex(fq(b[48:]))}
        elif len(b) == 192:
            out[k] = {"group": "G2", "x": [hex(fq(b[0:48])), hex(fq(b[48:96]))], "y": [hex(fq(b[96:144])), hex(fq(b[144:192]))]}
    json.dump({"source": SRC, "points": out}, open(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/verifier_crs_kat.json", "w"), indent=1)
    print(sorted(out))


if __name__ == "__main__":
    main()
