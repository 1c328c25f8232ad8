"""Host-side helpers of the drop-in package: config layer, seeding/sampler wrapper, distributed wrapper."""
