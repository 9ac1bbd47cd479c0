"""Synthetic netlist generators for the BASELINE.json configs.  Inputs only — no solver logic."""


def resistor_ladder(n: int, r: str = "1000") -> str:
    """circuits/resistor_ladder_{500,1000,5000}: V1 1 0 1; R{k} k k+1 1000; R{n} n 0 1000."""
    lines = ["% N-sized resistor ladder", "V1 1 0 1"]
    for k in range(1, n):
        lines.append(f"R{k} {k} {k + 1} {r}")
    lines.append(f"R{n} {n} 0 {r}")
    return "\n".join(lines) + "\n"


def rc_ladder(n: int, r: str = "1000", c: str = "1n") -> str:
    """BASELINE config 3: ladder + C{k} k 0 1n, order V, all R, all C (same index map as the resistor ladder)."""
    lines = ["* RC ladder", "V1 1 0 1 AC 1"]
    for k in range(1, n):
        lines.append(f"R{k} {k} {k + 1} {r}")
    lines.append(f"R{n} {n} 0 {r}")
    for k in range(1, n + 1):
        lines.append(f"C{k} {k} 0 {c}")
    return "\n".join(lines) + "\n"


def rlc_mesh(rows: int, cols: int, transient: bool = False) -> str:
    """BASELINE config 5 generator at arbitrary grid size."""
    L = ["* RLC mesh", "V1 src 0 dc 0 AC 1 PULSE(0 1 1u 1u 1u 1 1)", "Rs src n0_0 1"]
    k = 0
    for r in range(rows):
        for c in range(cols - 1):
            L.append(f"Rh{k} n{r}_{c} n{r}_{c + 1} 100"); k += 1
    k = 0
    for r in range(rows - 1):
        for c in range(cols):
            L.append(f"Lv{k} n{r}_{c} n{r + 1}_{c} 1u"); k += 1
    k = 0
    for r in range(rows):
        for c in range(cols):
            L.append(f"Cg{k} n{r}_{c} 0 1p"); k += 1
    L.append(f"RL n{rows - 1}_{cols - 1} 0 50")
    if transient:
        L.append(f"I1 0 n{rows // 2}_{cols // 2} 1m")
    return "\n".join(L) + "\n"


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def mc_factor(lane: int, j: int, nvar: int, seed: int = 0x6B72657473) -> float:
    """v = nominal*(0.95 + 0.10*u), u = (splitmix64(seed + nvar*lane + j) >> 11) * 2^-53; lane 0 nominal."""
    if lane == 0:
        return 1.0
    u = (splitmix64(seed + nvar * lane + j) >> 11) * (2.0 ** -53)
    return 0.95 + 0.10 * u
