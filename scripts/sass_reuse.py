import re, subprocess, sys
def analyze(path, kern="_ZN3eps20numerov_sweep_kernelILi2ELi8ELi32ELb0ELb0E"):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    m = re.search(r"Function : " + kern + r".*?(?=Function :|\Z)", txt, re.S)
    lines = [l for l in m.group(0).splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    ins = []
    for l in lines:
        body = l.split("*/", 1)[1].split(";")[0].strip()
        ins.append(body)
    good = bad = 0
    for i, b in enumerate(ins):
        if not b.startswith("DFMA"):
            continue
        ops = [o.strip() for o in b[4:].split(",")]
        srcs = ops[1:]
        regs = [re.sub(r"[-|]|\.reuse", "", o) for o in srcs]
        if not all(r.startswith("R") for r in regs) or len(set(regs)) < 3:
            continue  # immediate / constant operand or repeated register: <= 2 RF reads
        # previous FP64 instruction
        prev = ins[i - 1]
        pops = [o.strip() for o in prev.split(None, 1)[1].split(",")][1:] if " " in prev else []
        hit = False
        for slot, o in enumerate(pops):
            if ".reuse" in o and slot < len(regs) and re.sub(r"[-|]|\.reuse", "", o) == regs[slot]:
                hit = True
        if hit:
            good += 1
        else:
            bad += 1
    return good, bad
if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(p, "3-reg DFMA fed by reuse / not:", analyze(p))
