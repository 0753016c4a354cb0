from .input_stage import (ImageBatch, TwoViewInputStage, color_jitter_plan, color_jitter_u8, random_resized_crop_params,  # noqa: F401
                          resized_crop_u8, views_finalize)
