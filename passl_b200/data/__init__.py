from .input_stage import (ImageBatch, TwoViewInputStage, ViewRecipe, build_input_stage, color_jitter_plan, color_jitter_u8,  # noqa: F401
                          gaussian_blur_u8, gaussian_taps_fixed, grayscale_u8, random_resized_crop_params, resized_crop_u8,
                          views_finalize)
