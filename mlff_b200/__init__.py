"""mlff_b200 -- B200-native SO3krates energy+force hot path behind the C ABI in include/so3k.h.

Importing the package does not need a GPU; constructing an engine does (there is no CPU fallback).
"""
from .config import So3kratesConfig, md17_property_keys  # noqa: F401

__version__ = '0.1.0'


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import mlff_b200` stays cheap
    if name in ('So3krates', 'StackNet', 'init_stack_net', 'get_obs_and_force_fn', 'get_observable_fn',
                'get_energy_force_stress_fn', 'get_grad_observable_fn', 'get_obs_and_grad_obs_fn'):
        from . import nn
        return getattr(nn, name)
    if name in ('MLFFPotential', 'Graph'):
        from . import potential
        return getattr(potential, name)
    if name == 'mlffCalculator':
        from . import calculator
        return calculator.mlffCalculator
    if name in ('Trainer', 'TrainState', 'Optimizer'):
        from . import training
        return getattr(training, name)
    if name in ('evaluate_model', 'mae_metric', 'mse_metric', 'rmse_metric', 'r2_metric'):
        from . import inference
        return getattr(inference, name)
    if name == 'So3kEngine':
        from . import engine
        return engine.So3kEngine
    raise AttributeError(name)
