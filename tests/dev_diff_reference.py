"""dev tool (NOT collected by pytest; build container only: needs /root/reference): differential test of whole DMRG runs
against the LIVE unmodified reference on random small cases -- model (TFI, TFI with parity, XXZ-Sz, Hubbard N,Sz),
couplings, L, chi, one / two active sites, mixer, combine, matvec route ('combined' / 'split' incl. the identity-environment
shortcut / 'auto') -- on the numpy test double of the device library with NaN-poisoned uninitialised buffers.

    python tests/dev_diff_reference.py [seed] [n_cases]

End of round 1: seeds 1 and 2, 34 cases, all energies equal to 1e-14 and entropies to 1e-13.
"""
import sys, os, warnings, time
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
from tenpy_b200.algorithms import dmrg as mydmrg
import tenpy
from tenpy.models.tf_ising import TFIChain
from tenpy.models.spins import SpinChain
from tenpy.models.hubbard import FermiHubbardChain
from tenpy.networks.mps import MPS
from tenpy.algorithms import dmrg
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
n=int(sys.argv[2]) if len(sys.argv)>2 else 12
bad=0
for case in range(n):
    kind=rng.choice(['tfi','tfip','xxz','hub'])
    L=int(rng.choice([6,8,10])); chi=int(rng.choice([16,24,32]))
    active=int(rng.choice([1,2])); order=str(rng.choice(['combined','split','auto']))
    if kind in ('tfi','tfip'):
        g=float(rng.uniform(0.5,1.5)); cons=None if kind=='tfi' else 'parity'
        Mr=TFIChain(dict(L=L,J=1.,g=g,bc_MPS='finite',conserve=cons)); Mm=mym.TFIChain({'L':L,'J':1.,'g':g,'conserve':cons}); st=['up']*L
    elif kind=='xxz':
        jz=float(rng.uniform(0.3,1.5))
        Mr=SpinChain(dict(L=L,S=0.5,Jx=1.,Jy=1.,Jz=jz,bc_MPS='finite',conserve='Sz')); Mm=mym.SpinChain({'L':L,'Jx':1.,'Jy':1.,'Jz':jz,'conserve':'Sz'}); st=['up','down']*(L//2)
    else:
        L=6; U=float(rng.uniform(1.,6.))
        Mr=FermiHubbardChain(dict(L=L,t=1.,U=U,mu=0.,bc_MPS='finite',cons_N='N',cons_Sz='Sz')); Mm=mym.FermiHubbardChain({'L':L,'t':1.,'U':U,'mu':0.}); st=['up','down']*(L//2)
    mixer = True if (kind!='tfi' or active==1) else bool(rng.integers(0,2))
    opts=dict(mixer=mixer, mixer_params=dict(amplitude=1e-3,decay=2.,disable_after=8), max_E_err=1e-11, max_S_err=1e-8,
              trunc_params=dict(chi_max=chi, svd_min=1e-10), combine=bool(rng.integers(0,2)) if active==1 else True, max_sweeps=24, active_sites=active)
    psi_r=MPS.from_product_state(Mr.lat.mps_sites(), st, bc='finite')
    try:
        Er=dmrg.run(psi_r, Mr, dict(opts))['E']
    except Exception as e:
        print(case,'reference raised',type(e).__name__,'-> skipped'); continue
    psi_m=MyMPS.from_product_state(Mm.lat_sites, st)
    o2=dict(opts); o2['matvec_order']=order
    t0=time.time(); Em=mydmrg.run(psi_m, Mm, o2)['E']
    dS=np.max(np.abs(psi_m.entanglement_entropy()-psi_r.entanglement_entropy()))
    ok = abs(Em-Er) < 1e-9*max(1,abs(Er)) and dS < 1e-6
    print('%2d %-4s L=%2d chi=%2d sites=%d mixer=%s order=%-8s combine=%s  E_ref %.12f  dE %.1e  dS %.1e  %s' % (case,kind,L,chi,active,mixer,order,opts['combine'],Er,Em-Er,dS,'ok' if ok else 'MISMATCH'), flush=True)
    bad += (not ok)
print('mismatches', bad)
