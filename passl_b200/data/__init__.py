from .input_stage import ImageBatch, TwoViewInputStage, random_resized_crop_params, resized_crop_u8, views_finalize  # noqa: F401
