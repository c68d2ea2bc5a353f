-- Stub `spy` package.  The reference executes `require 'spy'` (pwc.lua:26, back2future.lua:27,
-- util.lua:14, donkey.lua:14) but never instantiates nn.ScaleBHWD, so an empty package suffices.
require 'nn'
return nn
