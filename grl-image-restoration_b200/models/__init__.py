"""Re-exports at the reference's dotted paths (models.networks.grl.GRL, models.common.*) so that a Hydra `_target_`
only needs the package prefix changed.  All implementations live in ..modules / ..geometry."""
