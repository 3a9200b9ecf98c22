from .radiance_field import (  # noqa: F401
    DensityField,
    RadianceField,
    build_density_field,
    build_radiance_field_from_cfg,
)
