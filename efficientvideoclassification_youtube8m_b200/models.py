"""Contains the base class for models (code_student_uniform/models.py:4-8)."""


class BaseModel(object):
    """Inherit from this class when implementing new models."""

    def create_model(self, unused_model_input, **unused_params):
        raise NotImplementedError()
