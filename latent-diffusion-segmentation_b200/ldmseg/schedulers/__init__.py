from .ddim_scheduler import DDIMNoiseScheduler

__all__ = ["DDIMNoiseScheduler"]
