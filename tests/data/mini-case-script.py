from utils import *
from samplers import *

#---- Parameters (a small driver in the style of the reference's case scripts, Python-2 syntax on purpose)
Niter = 120
Nchain = 12
N_warm_up = 40
R_thin = 1
N_save_chain0 = 5
dt = 1e-1
L_low = 5
L_high = 20

print "#---- mini case ----#"
title_str = "./mini/mini"
D = 60
q0 = np.zeros(D, dtype=np.float)
rho = 0.5
cov0 = np.diag(np.ones(D)) * (1-rho)
cov0 += rho
inv_cov0 = np.linalg.inv(cov0)

def V(q):
    return -normal_lnL(q, q0, cov0)

def dVdq(q):
    return np.dot(inv_cov0, (q-q0))

cov_start = np.diag(np.ones(D)) * 2
q_start = start_pts(q0, cov_start, Nchain)

HMC1 = HMC_sampler(D, V, dVdq, Niter=Niter, Nchain=Nchain, sampler_type="Random", L_low=L_low, \
                  L_high=L_high, dt=dt, thin_rate=R_thin, warm_up_num = N_warm_up)
HMC1.gen_sample(q_start, N_save_chain0 = N_save_chain0)
HMC1.compute_convergence_stats()
HMC1.plot_samples(title_prefix=title_str, savefig=True, show=False, plot_normal=True, q0=q0, cov0=cov0)
HMC1.make_movie(title_prefix=title_str, q0=q0, cov0=cov0, plot_cov=True, qmin=-4, qmax=4)

print "Random"
print "Total number of samples: %d" % ((HMC1.L_chain-1) * HMC1.Nchain)
print "Effective number per param: ", HMC1.n_eff_q
print "Ratio", HMC1.n_eff_q/((HMC1.L_chain-1) * HMC1.Nchain)
print "\n"
