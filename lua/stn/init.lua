-- Replacement `stn` package: only what Back2Future uses (nn.BilinearSamplerBHWD).  The reference's
-- init.lua:9-14 also requires the Affine* modules and stn.test, which nothing in Back2Future touches.
require 'nn'
require 'cutorch'
require('stn.BilinearSamplerBHWD')
return nn
