#!/usr/bin/env python
"""SASS listings of the hot kernels + mnemonic counts (what proves TMA bulk copies / no calls / no tensor cores):

    python tools/sass_summary.py            # writes profiles/r02_sass_{replay_flat,replay_sorted,env,replay_flat_L50,replay_hyb_L50}.txt

Reads the object files of the in-tree build (rl4mm_b200/_native/obj), i.e. exactly what liblobsim.so was linked from."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / "rl4mm_b200" / "_native" / "obj"
MNEMONICS = ["UBLKCP", "SYNCS", "CALL", "VOTE", "CREDUX", "REDUX", "SHFL", "LDS", "STS", "LDL", "STL", "LDG", "STG", "BAR", "WARPSYNC", "UTCMMA", "UTMALDG", "LDTM",
             "HMMA", "DFMA", "DMUL", "DADD", "MUFU"]


def functions(obj: Path):
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    cur, name = [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, cur
            name, cur = m.group(1), []
        elif name:
            cur.append(line)
    if name:
        yield name, cur


def demangle(name: str) -> str:
    return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name


def main():
    round_tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    targets = [("replay_flat", "fast0_0.o", lambda n: "k_replay_flat" in n), ("replay_sorted", "fast0_0.o", lambda n: "k_replay_fast" in n),
               ("env", "fast2_0.o", lambda n: "k_env_fast" in n and "Lb1ELb0" in n),
               ("replay_flat_L50", "fast3_0.o", lambda n: "k_replay_flat" in n),
               ("replay_hyb_L50", "fast4_0.o", lambda n: "k_replay_hyb" in n)]          # layouts.h entry 4 = 128/1024/64
    for tag, obj, pred in targets:
        for name, lines in functions(OBJ / obj):
            if not pred(name):
                continue
            instr = [l for l in lines if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
            counts = {m: sum(1 for l in instr if re.search(r"\b" + m + r"(\.|\b)", l)) for m in MNEMONICS}
            head = [f"# {demangle(name)}", f"# object: rl4mm_b200/_native/obj/{obj} (nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3)",
                    f"# {len(instr)} SASS instructions; " + ", ".join(f"{m} {c}" for m, c in counts.items()), ""]
            keep_listing = tag != "env"                     # the env kernel is 17 K instructions: counts only + its first lines
            body = instr if keep_listing else instr[:200] + ["        ... (listing truncated: regenerate with tools/sass_summary.py)"]
            path = ROOT / "profiles" / f"{round_tag}_sass_{tag}.txt"
            path.write_text("\n".join(head + [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in body]) + "\n")
            print(path.name, len(instr), {k: v for k, v in counts.items() if v})
            break


if __name__ == "__main__":
    main()
