"""Shared helpers for the tests: fixture loading and tape reconstruction."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

RANDOM_FIXTURES = ["random_case1a", "random_d10_thin3", "random_case3c_small", "random_case3d_small",
                   "random_case2c_small", "random_vecdt_covp"]
NUTS_FIXTURES = ["nuts_d2", "nuts_d10", "nuts_case3c_small"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def flat_tape(fx):
    """Structured per-chain tapes of a fixture -> flat ("p"/"i"/"u") tape in the reference's consumption order.

    For NUTS the interleaving of coins and uniforms is data dependent, so the flat order cannot be rebuilt from
    the arrays alone; `ChainTapeDraws` below serves each kind from its own per-chain queue instead."""
    Nchain, Niter = int(fx["Nchain"]), int(fx["Niter"])
    tape = []
    for m in range(Nchain):
        tape.append(("p", fx["p_tape"][m, 0]))
        for i in range(Niter):
            tape.append(("p", fx["p_tape"][m, i + 1]))
            tape.append(("i", int(fx["L_tape"][m, i])))
            tape.append(("u", float(fx["u_tape"][m, i])))
    return tape


class ChainTapeDraws(object):
    """Draw source fed by per-chain queues (p rows, ints, uniforms); chains are consumed one after another.
    A chain is finished once Niter+1 momenta were served, which is how the samplers consume them."""

    def __init__(self, p_tape, i_tape, u_tape):
        self.p_tape, self.i_tape, self.u_tape = p_tape, i_tape, u_tape
        self.m = 0
        self.np = 0
        self.ni = 0
        self.nu = 0
        self.per_chain = p_tape.shape[1]

    def p_sample(self):
        if self.np == self.per_chain:
            self.m += 1
            self.np = self.ni = self.nu = 0
        v = np.array(self.p_tape[self.m, self.np], dtype=float)
        self.np += 1
        return v

    def randint(self, low, high):
        v = int(self.i_tape[self.m, self.ni])
        self.ni += 1
        return v

    def random(self):
        v = float(self.u_tape[self.m, self.nu])
        self.nu += 1
        return v
