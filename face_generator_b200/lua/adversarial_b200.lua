-- adversarial_b200.lua -- drop-in for adversarial.lua's loop body (adversarial.lua:54-300): the whole
-- "1 D iteration + 1 G iteration" (batch assembly excluded) is ONE fg_train_step call.  Keeps the globals
-- train.lua sets up (OPT, OPTSTATE, CONFUSION, NN_UTILS, IMG_DIMENSIONS) and the ADVERSARIAL.train signature.
-- Delivered untested-by-execution (no LuaJIT/Torch7 in the build image); face_generator_b200/adversarial.py
-- is the executable mirror.
require 'torch'
local ffi = require 'ffi'
local F = require 'fg_ffi'
require 'b200'
local C = F.C

local adversarial = {}
adversarial.accs = {}

function adversarial.train(dataset, maxAccuracyD, accsInterval)
  EPOCH = EPOCH or 1
  local N_epoch = OPT.N_epoch
  if N_epoch <= 0 then N_epoch = dataset:size() end
  local dataBatchSize = OPT.batchSize / 2
  local ctx = b200.context(OPT.gpu, OPT.batchSize, IMG_DIMENSIONS[1])
  local hyper = b200.hyperFromOPT(OPT, OPTSTATE)
  hyper[0].D_maxAcc, hyper[0].accs_interval = maxAccuracyD, accsInterval
  local stats = ffi.new('fg_step_stats[1]')
  local time = sys.clock()
  local seed = (EPOCH - 1) * 1000000
  for t = 1, N_epoch, dataBatchSize do
    local thisBatchSize = math.min(OPT.batchSize, N_epoch - t + 1)
    if thisBatchSize < 4 then break end                      -- adversarial.lua:73-76
    thisBatchSize = thisBatchSize - thisBatchSize % 2        -- even batches only (SURVEY appendix 13)
    local half = thisBatchSize / 2
    -- (1.1) real half-batch (adversarial.lua:244-249)
    local real = torch.FloatTensor(half, IMG_DIMENSIONS[1], IMG_DIMENSIONS[2], IMG_DIMENSIONS[3])
    for i = 1, half do real[i] = dataset[math.random(dataset:size())] end
    -- (1.2)/(2) noise for the D-step fakes and for the G step (nn_utils.lua:35-39)
    local noiseD = NN_UTILS.createNoiseInputs(half)
    local noiseG = NN_UTILS.createNoiseInputs(thisBatchSize)
    seed = seed + 1
    F.check(C.fg_train_step(ctx, hyper, thisBatchSize, F.ptr(real), F.ptr(noiseD), F.ptr(noiseG), nil, nil, seed, stats), 'fg_train_step')
    local s = stats[0]
    -- feed optim.ConfusionMatrix exactly like adversarial.lua:112-117 (rows = predicted class, cols = target)
    CONFUSION.mat[2][2] = CONFUSION.mat[2][2] + s.conf[0]
    CONFUSION.mat[1][2] = CONFUSION.mat[1][2] + s.conf[1]
    CONFUSION.mat[2][1] = CONFUSION.mat[2][1] + s.conf[2]
    CONFUSION.mat[1][1] = CONFUSION.mat[1][1] + s.conf[3]
    OPTSTATE.adam.D.t, OPTSTATE.adam.G.t = s.t_D, s.t_G
    xlua.progress(t + thisBatchSize, N_epoch)
  end
  time = sys.clock() - time
  print(string.format("<trainer> time to learn 1 sample = %f ms", 1000 * time / N_epoch))
  print(CONFUSION)
  CONFUSION:zero()
  EPOCH = EPOCH + 1
end

return adversarial
