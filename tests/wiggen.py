"""Random wiggle text for the halWiggleLiftover tests (fixedStep / variableStep, spans, steps, odd values)."""


def random_wig(rng, seqs, sections=(2, 6), max_lines=400, disorder=0.0):
    """seqs: [(name, start, length)].  disorder > 0 sometimes emits sections that step backwards inside a sequence."""
    out = []
    for _ in range(rng.randint(*sections)):
        nm, _, ln = rng.choice(seqs)
        if ln < 60:
            continue
        kind = rng.choice(["fixed", "fixedspan", "var", "varspan"])
        if kind.startswith("fixed"):
            step = rng.choice([1, 1, 1, 3, 10])
            span = rng.choice([1, 2, 5]) if kind == "fixedspan" else None
            if span and span > step and rng.random() < 0.9:  # overlapping lines are "Coordinate out of order" inside a batch
                step = span + rng.choice([0, 0, 3])
            n = rng.randint(1, max(1, min(max_lines, (ln - 10) // step)))
            start = rng.randint(1, max(1, ln - n * step - (span or 1)))
            out.append(f"fixedStep chrom={nm} start={start} step={step}" + (f" span={span}" if span else ""))
            for _ in range(n):
                v = rng.choice([rng.random() * 100 - 20, rng.randint(-3, 50), 1e-7 * rng.random(), 123456789.0 * rng.random()])
                out.append(f"{v:.6g}")
        else:
            span = rng.choice([2, 4, 30]) if kind == "varspan" else None
            out.append(f"variableStep chrom={nm}" + (f" span={span}" if span else ""))
            pos = rng.randint(1, ln // 2)
            for _ in range(rng.randint(1, max_lines // 2)):
                pos += (span or 1) + rng.randint(0, 40)
                if disorder and rng.random() < disorder:
                    pos = max(1, pos - rng.randint(1, 300))
                if pos + (span or 1) >= ln:
                    break
                out.append(f"{pos} {rng.random() * 10:.3f}")
    return "\n".join(out) + "\n"
