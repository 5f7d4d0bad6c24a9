import re, numpy as np, mpmath as mp
mp.mp.dps = 50
src = open('/root/reference/src/cpp/AnisoVariables.h').read()
def ref_table(name):
    body = re.search(name + r"\[pbar_pts\]\s*=\s*\{(.*?)\};", src, flags=re.S).group(1)
    return np.array([float(x) for x in body.replace('\n',' ').split(',')])
def genlag(n, a):
    # roots of generalized Laguerre L_n^{(a)} via mpmath polyroots; weights w_i = Gamma(n+a+1) x_i / (n! (n+1)^2 [L_{n+1}^{(a)}(x_i)]^2)
    coeffs = [mp.binomial(n + a, n - k) * (-1)**k / mp.factorial(k) for k in range(n + 1)]  # ascending powers
    roots = mp.polyroots(list(reversed(coeffs)), maxsteps=2000, extraprec=400)
    roots = sorted([mp.re(r) for r in roots])
    ws = []
    for x in roots:
        L = mp.laguerre(n + 1, a, x)
        ws.append(mp.gamma(n + a + 1) * x / (mp.factorial(n) * (n + 1)**2 * L**2))
    return roots, ws
out = []
for a in (1, 2, 3):
    r, w = genlag(16, a)
    rr, rw = ref_table(f"pbar_root_a{a}"), ref_table(f"pbar_weight_a{a}")
    r64 = np.array([float(x) for x in r]); w64 = np.array([float(x) for x in w])
    print(a, "max rel diff roots", np.abs(r64/rr-1).max(), "weights", np.abs(w64/rw-1).max())
    out.append((a, r, w))
with open("/tmp/gl16_tables.txt", "w") as f:  # pasted into is3d2_b200/csrc/aniso_gl16.inc
    for a, r, w in out:
        f.write(f"// alpha = {a}\n")
        f.write("{" + ", ".join(mp.nstr(x, 20) for x in r) + "},\n")
        f.write("{" + ", ".join(mp.nstr(x, 20) for x in w) + "},\n")
