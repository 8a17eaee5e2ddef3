"""Import shim with the reference's module paths (`utils.lora_modules`, `utils.models`, `utils.noise_layers.noiser`): put the
`compat/` directory ahead of the reference's own `utils/` on sys.path and the reference's training / evaluation scripts
pick up the B200 kernels unchanged."""
