import sys, tempfile
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cases, harness
name=sys.argv[1]
surf, ref = harness.load_golden(name)
with tempfile.TemporaryDirectory() as tmp:
    with harness.open_session(tmp, cases.SPECTRA_CASES[name], surf) as h:
        got, st = h.abi_spectra()
rel=np.abs(got-ref)/np.abs(ref)
peak=np.abs(ref).reshape(ref.shape[0],-1).max(axis=1)
idx=np.argsort(rel.ravel())[::-1][:8]
for i in idx:
    j=np.unravel_index(i, ref.shape)
    print(j, "rel %.2e"%rel[j], "ref %.3e"%ref[j], "ref/peak %.2e"%(abs(ref[j])/peak[j[0]]))
print("median rel", np.median(rel), "95pct", np.percentile(rel,95))
