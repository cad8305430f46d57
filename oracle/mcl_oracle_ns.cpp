// placeholder until the NS oracle lands
