"""Box space with the attributes the planner reads (.low, .high, .shape); mirrors learning_to_adapt/spaces/box.py:5-45."""
import numpy as np


class Box(object):
    def __init__(self, low, high, shape=None):
        if shape is None:
            low, high = np.asarray(low, np.float64), np.asarray(high, np.float64)
            assert low.shape == high.shape
            self.low, self.high = low, high
        else:
            assert np.isscalar(low) and np.isscalar(high)
            self.low = low + np.zeros(shape)
            self.high = high + np.zeros(shape)

    def sample(self):
        return np.random.uniform(low=self.low, high=self.high, size=self.low.shape)

    def sample_n(self, n):
        return np.random.uniform(low=self.low, high=self.high, size=(n,) + self.low.shape)

    def contains(self, x):
        return x.shape == self.shape and (x >= self.low).all() and (x <= self.high).all()

    @property
    def shape(self):
        return self.low.shape

    @property
    def flat_dim(self):
        return int(np.prod(self.low.shape))

    @property
    def bounds(self):
        return self.low, self.high
