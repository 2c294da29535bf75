from .panda_tasks import PandaFlipEnv, PandaPickAndPlaceEnv, PandaPushEnv, PandaReachEnv, PandaSlideEnv, PandaStackEnv

__all__ = ["PandaReachEnv", "PandaPushEnv", "PandaSlideEnv", "PandaPickAndPlaceEnv", "PandaStackEnv", "PandaFlipEnv"]
