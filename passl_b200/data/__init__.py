from .input_stage import (DeviceAugmentedTwoViews, ImageBatch, SyntheticDecodedImages, TwoViewInputStage, ViewRecipe,  # noqa: F401
                          build_input_stage, color_jitter_plan, color_jitter_u8, gaussian_blur_u8, gaussian_taps_fixed, grayscale_u8,
                          random_resized_crop_params, resized_crop_u8, views_finalize)
