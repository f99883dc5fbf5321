"""Drop-in replacement for the reference's `TPAMI_attack` module (adaptive ENS-I2V, TPAMI'24).

`AENS_I2V_MF(model_name_lists, depths, step_size, momentum=0, coef_CE=False, epsilon=16/255,
steps=60)(videos, labels, video_names) -> (adv, used_time, cost_saved)` exactly as reference
TPAMI_attack.py:141-320, including its quirks:
  * `coeffs` has 2 entries per model (TPAMI_attack.py:165) and persists across calls;
    here a depth list of any length is accepted as long as the total is 2 * len(models), and a
    mismatch raises ValueError instead of failing inside a broadcast.
  * with a list of depths SqueezeNet hooks the whole Fire module (TPAMI_attack.py:195-198, D9).
  * `weights` records the coefficient vector of every step (TPAMI_attack.py:266).
"""
import time

import numpy as np
import torch

from i2v_b200 import attack_loop, backbones, engines
from image_attacks import Attack, get_model, get_models

__all__ = ["Attack", "get_model", "get_models", "AENS_I2V_MF"]


class AENS_I2V_MF(Attack):
    """The adaptive I2V attack with multiple models and layers.

    Parameters:
        model_name_lists: surrogate image model names, e.g. ['resnet', 'vgg', 'squeezenet', 'alexnet']
        depths: layers used in each model, e.g. {'resnet':[2,3], 'vgg':[2,3], ...}
        step_size: the learning rate.
    Return:
        image_inps: adversarial video; used_time: seconds spent in the loop; cost_saved: cost per step
    """

    def __init__(self, model_name_lists, depths, step_size, momentum=0, coef_CE=False, epsilon=16 / 255, steps=60,
                 *, engine=None, placement=None):
        super(AENS_I2V_MF, self).__init__("AENS_I2V_MF")
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = step_size
        self.loss_info = {}
        self.depths = depths
        self.momentum = momentum
        self.coef_CE = coef_CE
        self.model_names = model_name_lists
        # placement='ensemble' (extension): one backbone per GPU under torch.distributed — this rank builds only its
        # members; dcost/dtrue_image and the cosine rows are summed over the ensemble group every step (i2v_b200/dist.py)
        self._plan = None
        layers_per_model = [len(depths[n]) if isinstance(depths[n], (list, tuple)) else 1 for n in model_name_lists]
        if placement == "ensemble":
            from i2v_b200 import dist as D
            self._plan = D.EnsemblePlan(model_name_lists, layers_per_model)
            mine = [model_name_lists[i] for i in self._plan.members]
        elif placement is None:
            mine = list(model_name_lists)
        else:
            raise ValueError("placement must be None or 'ensemble', got %r" % (placement,))
        self.models = get_models(mine)
        self._engines = [engines.make_engine(m, n, depths[n], engine) for m, n in zip(self.models, mine)]
        n_layers = sum(layers_per_model)
        if n_layers != 2 * len(model_name_lists):
            raise ValueError("AENS_I2V_MF keeps 2 coefficients per model (reference TPAMI_attack.py:165): "
                             "%d models need %d hooked layers in total, depths give %d"
                             % (len(model_name_lists), 2 * len(model_name_lists), n_layers))
        device = torch.device("cuda", torch.cuda.current_device())
        self.coeffs = torch.ones(len(model_name_lists) * 2, device=device)
        self.weights = []

    def forward(self, videos, labels, video_names):
        begin = time.time()
        extra = {}
        if self._plan is not None:
            extra = dict(reduce_hook=self._plan.hook(), layer_offsets=self._plan.layer_offsets,
                         n_layers_total=self._plan.n_layers_total)
        res = attack_loop.run_image_guided(self._engines, videos, self.epsilon, self.steps, self.step_size,
                                           adaptive=True, coeffs=self.coeffs, momentum=self.momentum,
                                           coef_CE=self.coef_CE, cache=self.__dict__.setdefault("_run_cache", {}), **extra)
        self.weights = [w.copy() for w in res.weights] if res.weights is not None else []
        cost_saved = np.zeros(self.steps)
        cost_saved[:] = res.cost
        attack_loop.record_loss_info(self.loss_info, video_names, res.cost)
        # the loop is asynchronous; the cost log copy above is its synchronisation point, so
        # `used_time` is a true wall-clock of the attack (the reference's is un-synchronised)
        used_time = time.time() - begin
        return res.adv, used_time, cost_saved
