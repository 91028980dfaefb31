"""dev tool (NOT collected by pytest; build container only: needs /root/reference): TEBD / QR based TEBD (imaginary time:
sweeps and brick-wall steps of order 1, 2, 4) against the LIVE unmodified reference on random small cases, on the numpy test
double with NaN-poisoned uninitialised buffers.  End of round 1: 8 cases, bond energies / entropies equal to 1e-12, chi exact.

    python tests/dev_diff_reference_tebd.py
"""
import sys, os, warnings
os.environ['TENPY_NO_CYTHON']='1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.environ.get('TENPY_REFERENCE', '/root/reference'))
warnings.simplefilter('ignore')
import numpy as np, torch
from tenpy_b200 import backend
from fake_device import FakeDeviceLib
backend.use_library(FakeDeviceLib())
backend.empty = lambda n: torch.full((int(n),), float('nan'), dtype=torch.float64)
from tenpy_b200 import models as mym
from tenpy_b200.networks.mps import MPS as MyMPS
from tenpy_b200.algorithms import tebd as mytebd
from tenpy.models.tf_ising import TFIChain
from tenpy.models.spins import SpinChain
from tenpy.networks.mps import MPS
from tenpy.algorithms import tebd
rng=np.random.default_rng(4); bad=0
for case in range(8):
    kind=rng.choice(['tfi','xxz']); L=int(rng.choice([6,8,10])); chi=int(rng.choice([8,16,24])); qr=bool(rng.integers(0,2))
    order=int(rng.choice([1,2,4])); dt=float(rng.choice([0.02,0.05])); imag_sweeps=bool(rng.integers(0,2))
    if kind=='tfi':
        g=float(rng.uniform(0.6,1.4)); Mr=TFIChain(dict(L=L,J=1.,g=g,bc_MPS='finite',conserve=None)); Mm=mym.TFIChain({'L':L,'J':1.,'g':g,'conserve':None}); st=['up']*L
    else:
        jz=float(rng.uniform(0.3,1.5)); Mr=SpinChain(dict(L=L,S=0.5,Jx=1.,Jy=1.,Jz=jz,hz=0.,bc_MPS='finite',conserve='Sz')); Mm=mym.SpinChain({'L':L,'Jx':1.,'Jy':1.,'Jz':jz,'conserve':'Sz'}); st=['up','down']*(L//2)
    opts={'trunc_params':{'chi_max':chi,'svd_min':1e-8}}
    if qr: opts.update(cbe_expand=0.1, cbe_expand_0=0.5, cbe_min_block_increase=2, compute_err=True)
    pr=MPS.from_product_state(Mr.lat.mps_sites(), st, bc='finite'); pm=MyMPS.from_product_state(Mm.lat_sites, st)
    er=(tebd.QRBasedTEBDEngine if qr else tebd.TEBDEngine)(pr, Mr, dict(opts)); em=(mytebd.QRBasedTEBDEngine if qr else mytebd.TEBDEngine)(pm, Mm, dict(opts))
    if imag_sweeps:
        er.calc_U(2, dt, type_evo='imag'); er.update_imag(12, call_canonical_form=False)
        em.calc_U(2, dt, type_evo='imag'); em.update_imag(12)
    else:
        er.calc_U(order, dt, type_evo='imag'); er.evolve(5, dt)
        em.calc_U(order, dt, type_evo='imag'); em.evolve(5, dt)
    dE=np.max(np.abs(np.asarray(Mr.bond_energies(pr))-Mm.bond_energies(pm))); dS=np.max(np.abs(pr.entanglement_entropy()-pm.entanglement_entropy()))
    ok = dE<1e-9 and dS<1e-8 and list(pr.chi)==list(pm.chi) and abs(pr.norm-pm.norm)<1e-9*abs(pr.norm)
    print(case,kind,'L',L,'chi',chi,'qr',qr,'sweeps' if imag_sweeps else 'order %d'%order,'dE %.1e dS %.1e'%(dE,dS),'chi',max(pm.chi),'ok' if ok else 'MISMATCH'); bad+=(not ok)
print('mismatches',bad)
