"""Prove that a source change left the MEASURED kernels untouched: compare the instruction streams of
every kernel in two builds of libsimilarity_transform.so.

    cuobjdump -sass eigen_value_b200/libsimilarity_transform.so > /tmp/before.sass    # before the change
    ... edit, rebuild ...
    cuobjdump -sass eigen_value_b200/libsimilarity_transform.so > /tmp/after.sass
    python tools/sass_diff.py /tmp/before.sass /tmp/after.sass [--append-arg Li0E]

Instruction text is compared (addresses and encodings are ignored).  `--append-arg` maps the old
mangled names to new ones when a trailing template argument was added to the round kernels (e.g.
`Li0E` for `int STOP = 0`), so that `round_loop_kernel<4,0,512>` is compared with
`round_loop_kernel<4,0,512,0>`.  Exit code 1 if any kernel of the first build is missing or differs.

Used in round 1 when the stop-test and storage-type template parameters were added without GPU time
left to re-measure (`--append-arg Li0E,Li0Ef`): every round-loop kernel of the measured build is
instruction-identical in the new one; only the standalone find_max / stop kernels changed, on purpose.
"""
import argparse
import re
import sys


def split(path):
    funcs, cur = {}, None
    for line in open(path):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur is not None:
            m2 = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", line)
            if m2:
                funcs[cur].append(m2.group(1).strip())
    return funcs


def digest(funcs):
    import hashlib
    return {name: {"instructions": len(body), "sha256": hashlib.sha256("\n".join(body).encode()).hexdigest()}
            for name, body in funcs.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("before", nargs="?")
    ap.add_argument("after", nargs="?")
    ap.add_argument("--refresh", action="store_true",
                    help="disassemble the CURRENT eigen_value_b200/libsimilarity_transform.so and write the digest of its round-loop "
                         "kernels to profiles/r2_measured_kernels_sass_digest.json: run this in the tree a GPU measurement is "
                         "taken from, and commit the file together with the numbers")
    ap.add_argument("--write-digest", default=None,
                    help="write {kernel: sha256 of its instruction stream} of `before` to this JSON file and exit "
                         "(refreshes profiles/r1_measured_kernels_sass_digest.json after a re-measurement)")
    ap.add_argument("--append-arg", default=None,
                    help="mangled template argument(s) appended to kernels taking RoundParams, e.g. Li0E; "
                         "several alternatives separated by commas (Li0E,Li0Ef)")
    args = ap.parse_args()
    if args.refresh:
        import json, os, shutil, subprocess, tempfile
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
        with tempfile.NamedTemporaryFile("w", suffix=".sass") as f:
            subprocess.run([cuobjdump, "-sass", os.path.join(root, "eigen_value_b200", "libsimilarity_transform.so")], stdout=f, check=True)
            f.flush()
            d = {k: v for k, v in digest(split(f.name)).items() if "round_loop" in k}
        with open(os.path.join(root, "profiles", "r2_measured_kernels_sass_digest.json"), "w") as f:
            json.dump({"kernels": d}, f, indent=1, sort_keys=True)
        print(f"{len(d)} round-loop kernels")
        return 0
    if args.write_digest:
        import json
        with open(args.write_digest, "w") as f:
            json.dump({"kernels": digest(split(args.before))}, f, indent=1, sort_keys=True)
        return 0
    if not args.after:
        ap.error("two SASS dumps are needed for a comparison")
    b, a = split(args.before), split(args.after)
    bad = 0
    for name, body in b.items():
        cands = [name]
        if args.append_arg and "RoundParams" in name:
            for extra in args.append_arg.split(","):
                cands.append(name.replace("EEvNS_11RoundParamsE", extra + "EEvNS_11RoundParamsE"))
        found = next((c for c in cands if c in a), None)
        if found is None:
            print("MISSING in second build:", name)
            bad += 1
        elif a[found] != body:
            print(f"DIFF ({len(body)} vs {len(a[found])} instructions): {name}")
            bad += 1
    print(f"kernels compared: {len(b)}; missing or different: {bad}; kernels only in the second build: "
          f"{len(a) - (len(b) - bad)}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
