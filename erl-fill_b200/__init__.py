"""erlfill_b200 — B200-native (sm_100a) implementation of ERL-Fill's batched rollout-and-update
hot path behind the reference's env/agent API.  Import as `import erlfill_b200` (the directory is
named erl-fill_b200/; the repo-root module erlfill_b200.py aliases it).

The compute lives in csrc/*.cu behind the C ABI of include/erlfill.h (lib/liberlfill.so); this
package is the thin Python host side.  There is no CPU fallback.
"""
from . import _lib, conditions  # noqa: F401
from ._lib import ErlfillError, SO_PATH  # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so that ABI checks work without touching CUDA
    if name in ("VecFillEnv", "WeightEnv", "MultiConditionWeightEnv"):
        from . import envs
        return getattr(envs, name)
    if name == "VecERDDPG":
        from . import agents
        return agents.VecERDDPG
    if name == "VecEmotionSAC":
        from . import esac
        return esac.VecEmotionSAC
    if name in ("VecTD3", "VecSAC"):
        from . import td3sac
        return getattr(td3sac, name)
    if name in ("DDPGAgent", "EmotionModule", "EmotionModuleNone", "TD3Agent", "SACAgent", "EmotionSACAgent", "reset_actor_scaling"):
        from . import dropin
        return getattr(dropin, name)
    if name in ("init", "td3sac", "nets", "ddpg", "agents", "envs", "dropin", "loops", "esac", "parallel"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
