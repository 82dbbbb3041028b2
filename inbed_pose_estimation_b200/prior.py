"""Max-mixture GMM pose prior (reference smplify/prior.py:102-196, merged path).

The constants are derived exactly as the reference does (numpy, float64 determinants,
fp32 buffers); the evaluation and its gradient run inside the CUDA fit kernel
(csrc/fit_tile.cuh ph_prior_*).  Only the merged negative log-likelihood the reference
actually uses (use_merged=True, prior.py:227-229) is provided.
"""
import os
import pickle
import sys

import numpy as np
import torch


def gmm_constants(gmm, dtype=np.float32):
    """means [8,69], precisions [8,69,69], nll_weights [8] from a gmm dict or sklearn GMM
    (reference prior.py:130-160)."""
    if isinstance(gmm, dict):
        means, covars, weights = gmm['means'], gmm['covars'], gmm['weights']
    elif 'sklearn.mixture.gmm.GMM' in str(type(gmm)):
        means, covars, weights = gmm.means_, gmm.covars_, gmm.weights_
    else:
        raise TypeError('Unknown type for the prior: {}'.format(type(gmm)))
    means32 = np.asarray(means).astype(dtype)
    covs32 = np.asarray(covars).astype(dtype)
    precisions = np.stack([np.linalg.inv(c) for c in covs32]).astype(dtype)
    sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in np.asarray(covars)])
    const = (2 * np.pi) ** (69 / 2.)
    nll_weights = np.asarray(np.asarray(weights) / (const * (sqrdets / sqrdets.min()))).astype(dtype)
    return {'means': means32, 'precisions': precisions, 'nll_weights': nll_weights,
            'weights': np.asarray(weights).astype(dtype)}


class MaxMixturePrior(torch.nn.Module):
    """Holder of the GMM constants with the reference's constructor signature.

    ``forward(pose, betas)`` evaluates the merged negative log-likelihood [B] with plain torch
    ops; it exists for API compatibility (nothing on the hot path calls it - SMPLify passes
    the constants to the CUDA library instead)."""

    def __init__(self, prior_folder='prior', num_gaussians=6, dtype=torch.float32, epsilon=1e-16,
                 use_merged=True, **kwargs):
        super(MaxMixturePrior, self).__init__()
        if dtype != torch.float32:
            print('Unknown float type {}, exiting!'.format(dtype))
            sys.exit(-1)
        if not use_merged:
            raise NotImplementedError('only the merged log-likelihood (the path SMPLify uses) is provided')
        self.num_gaussians = num_gaussians
        self.epsilon = epsilon
        self.use_merged = use_merged
        full_gmm_fn = os.path.join(prior_folder, 'gmm_{:02d}.pkl'.format(num_gaussians))
        if not os.path.exists(full_gmm_fn):
            print('The path to the mixture prior "{}"'.format(full_gmm_fn) + ' does not exist, exiting!')
            sys.exit(-1)
        with open(full_gmm_fn, 'rb') as f:
            gmm = pickle.load(f, encoding='latin1')
        self._init_from_gmm(gmm)

    @classmethod
    def from_gmm(cls, gmm, num_gaussians=8):
        """Build the prior from an in-memory gmm dict (tests and benchmarks: no file on disk)."""
        self = cls.__new__(cls)
        torch.nn.Module.__init__(self)
        self.num_gaussians, self.epsilon, self.use_merged = num_gaussians, 1e-16, True
        self._init_from_gmm(gmm)
        return self

    def _init_from_gmm(self, gmm):
        try:
            consts = gmm_constants(gmm)
        except TypeError as e:
            print(str(e) + ', exiting!')
            sys.exit(-1)
        self.register_buffer('means', torch.tensor(consts['means']))
        self.register_buffer('precisions', torch.tensor(consts['precisions']))
        self.register_buffer('nll_weights', torch.tensor(consts['nll_weights']).unsqueeze(0))
        self.register_buffer('weights', torch.tensor(consts['weights']).unsqueeze(0))
        self.random_var_dim = self.means.shape[1]

    def native_constants(self):
        return {'means': self.means.detach().cpu().numpy(), 'precisions': self.precisions.detach().cpu().numpy(),
                'nll_weights': self.nll_weights.detach().cpu().numpy().reshape(-1)}

    def get_mean(self):
        return torch.matmul(self.weights, self.means)

    def merged_log_likelihood(self, pose, betas=None):
        diff = pose.unsqueeze(1) - self.means
        quad = (torch.einsum('mij,bmj->bmi', self.precisions, diff) * diff).sum(-1)
        return torch.min(0.5 * quad - torch.log(self.nll_weights), dim=1)[0]

    def forward(self, pose, betas=None):
        return self.merged_log_likelihood(pose, betas)
